// OFDM/TDL link kernels for Nr=3, Nt=2
#define B200_NR 3
#define B200_NT 2
#include "ofdm_tdl_inst.cuh"
