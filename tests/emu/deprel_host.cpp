// TEST INFRASTRUCTURE -- csrc/deprel.cu compiled for the host (see cuda_runtime.h in this directory); exports the same
// extern "C" entry points as libgptb200.so does for K10, taking HOST pointers.  Built by tests/emu/emu_build.py.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

unsigned long long g_gpt_launches = 0;
int g_gpt_pdl = 0;

namespace {
alignas(16) thread_local float sm[160 * 1024 / 4];   // the kernels' `extern __shared__ float sm[]`: one block runs at a time
}

#include "../../gcn_over_pruned_trees_b200/csrc/deprel.cu"
