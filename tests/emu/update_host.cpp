// TEST INFRASTRUCTURE -- csrc/update.cu (K7) compiled for the host (see cuda_runtime.h in this directory); exports
// gpt_update_partials / gpt_update_sqnorm / gpt_update_apply taking HOST pointers.  Built by tests/emu/emu_build.py.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

#include "../../gcn_over_pruned_trees_b200/csrc/update.cu"
