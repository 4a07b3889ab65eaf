// OFDM/TDL link kernels for Nr=4, Nt=3
#define B200_NR 4
#define B200_NT 3
#include "ofdm_tdl_inst.cuh"
