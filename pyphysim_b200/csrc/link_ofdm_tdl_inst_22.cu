// OFDM/TDL link kernels for Nr=2, Nt=2
#define B200_NR 2
#define B200_NT 2
#include "ofdm_tdl_inst.cuh"
