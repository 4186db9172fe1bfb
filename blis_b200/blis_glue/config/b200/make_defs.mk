#
# make_defs.mk for the b200 sub-configuration (pattern: config/generic/make_defs.mk).
# Host code is plain C99; the CUDA engine is linked in as libblis_b200.so.
#
THIS_CONFIG    := b200

CPPROCFLAGS    := -I$(B200_ROOT)/include
CMISCFLAGS     :=
CPICFLAGS      := -fPIC
CWARNFLAGS     :=

ifneq ($(DEBUG_TYPE),off)
CDBGFLAGS      := -g
endif

ifeq ($(DEBUG_TYPE),noopt)
COPTFLAGS      := -O0
else
COPTFLAGS      := -O2
endif

CKOPTFLAGS     := $(COPTFLAGS) -O3
CKVECFLAGS     :=
CROPTFLAGS     := $(CKOPTFLAGS)
CRVECFLAGS     := $(CKVECFLAGS)

# the engine (built by `python -m blis_b200.build` with nvcc for sm_100a)
LDFLAGS        += -L$(B200_ROOT)/blis_b200 -lblis_b200 -Wl,-rpath,$(B200_ROOT)/blis_b200

# Store all of the variables here to new variables containing the
# configuration name.
$(eval $(call store-make-defs,$(THIS_CONFIG)))
