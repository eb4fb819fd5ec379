// TEST INFRASTRUCTURE -- csrc/gemm_simt.cu (K3, fp32 FFMA mode) compiled for the host (see cuda_runtime.h in this
// directory); exports gpt_linear_{fwd,dgrad,wgrad,wgrad_acc,wgrad_rows}_f32 taking HOST pointers.  Built by
// tests/emu/emu_build.py.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

#include "../../gcn_over_pruned_trees_b200/csrc/gemm_simt.cu"
