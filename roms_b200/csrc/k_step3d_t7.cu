// roms_b200/csrc/k_step3d_t7.cu -- experimental variant of the production step3d_t kernel, built from the SAME source with
// S3T_EXP=1 (see the top of k_step3d_t6.cu); opt-in at run time: ROMS_B200_STEP3D_T_V7=1.  Bit-identical to the oracle on the
// CPU emulation (tests/test_emu.py); not yet timed on hardware.
#define S3T_EXP 1
#define step3d_t_v6_kernel step3d_t_v7_kernel
#define k_step3d_t_v6 k_step3d_t_v7
#include "k_step3d_t6.cu"
