// relu edition of the tcgen05 MLP kernels: see the note at the top of mlp_tc.cu
#define B200RL_TC_ACT 2   /* B200RL_ACT_RELU */
#include "mlp_tc.cu"
