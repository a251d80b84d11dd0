// Mixed-radix frame lengths of the small-frame decoder (decs.cu: decs_kernel), 32-bit arithmetic.
#define FRAD_TU_MIXED 32
#include "decs.cu"
