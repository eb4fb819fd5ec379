// TEST INFRASTRUCTURE -- csrc/batch.cu (K9) compiled for the host (see cuda_runtime.h in this directory); exports
// gpt_build_batch taking HOST pointers.  Built by tests/emu/emu_build.py into the same library as deprel_host.cpp.
#define GPT_HOST_EMULATION 1
#include "cuda_runtime.h"

#include "../../gcn_over_pruned_trees_b200/csrc/batch.cu"
