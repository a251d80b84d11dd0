// Mixed-radix frame lengths of the small-frame encoder (enc2048.cu: encs_kernel), 32-bit arithmetic.
#define FRAD_TU_MIXED 32
#include "enc2048.cu"
