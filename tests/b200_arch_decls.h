/* What bli_arch_config.h would declare once BLIS_CONFIG_B200 is added to it
   (frame/include/bli_arch_config.h:79,211); used only to syntax-check
   config/b200/bli_cntx_init_b200.c against the unmodified reference headers. */
#include "blis.h"
void bli_cntx_init_b200( cntx_t* cntx );
void bli_cntx_init_b200_ref( cntx_t* cntx );
