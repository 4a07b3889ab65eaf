// OFDM/TDL link kernels for Nr=4, Nt=1
#define B200_NR 4
#define B200_NT 1
#include "ofdm_tdl_inst.cuh"
