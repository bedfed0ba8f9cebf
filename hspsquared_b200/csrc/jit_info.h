// jit_info.h -- what a kernel built for one batch (hspsquared_b200/jit.py) says about itself: the 16 x u64 global `<entry>_info`
// next to its entry point.  hsp2g_set_kernel_image (hsp2g_api.cu) reads it and launches the kernel only while everything that
// was fixed at compile time still equals the batch's configuration.
#pragma once
#define HSP2G_JIT_MAGIC 0x48535032474a4954ull /* "HSP2GJIT" */
enum { JI_MAGIC, JI_LAUNCH_BYTES, JI_SMEM, JI_OP, JI_ACT, JI_HOURLY, JI_FMASK, JI_S0, JI_S1, JI_S2, JI_BITS, JI_WORD, JI_BLOCK, JI_SBROWS, JI_MASK, JI_OPTS, JI_N = 16 };
enum { JB_F32OUT = 1, JB_PLAIN = 2, JB_F32IN = 4, JB_SHARED = 8, JB_WORD = 16, JB_IOSTATIC = 32, JB_LAYSTATIC = 64, JB_QUAL = 128, JB_OPTS = 256, JB_METTBL = 512, JB_LINKED = 1024, JB_LOCKSTEP = 2048 };
