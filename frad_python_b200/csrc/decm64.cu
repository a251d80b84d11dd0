// Mixed-radix frame lengths of the small-frame decoder (decs.cu: decs_kernel), 64-bit arithmetic.
#define FRAD_TU_MIXED 64
#include "decs.cu"
