// OFDM/TDL link kernels for Nr=4, Nt=4
#define B200_NR 4
#define B200_NT 4
#include "ofdm_tdl_inst.cuh"
