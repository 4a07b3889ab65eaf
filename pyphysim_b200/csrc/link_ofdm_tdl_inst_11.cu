// OFDM/TDL link kernels for Nr=1, Nt=1
#define B200_NR 1
#define B200_NT 1
#include "ofdm_tdl_inst.cuh"
