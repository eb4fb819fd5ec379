"""TEST INFRASTRUCTURE -- build tests/emu/_build/libgpt_emu.so: csrc/{deprel,prune_csr,aggregate,pool3,gemm_simt,embed,update,batch}.cu compiled by g++ against the host
stand-in for the CUDA runtime in this directory (one fiber per CUDA thread; see cuda_runtime.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_build', 'libgpt_emu.so')
CSRC = os.path.join(HERE, '..', '..', 'gcn_over_pruned_trees_b200', 'csrc')
SRCS = [os.path.join(HERE, f) for f in ('deprel_host.cpp', 'prune_host.cpp', 'batch_host.cpp', 'pool_host.cpp', 'gemm_host.cpp', 'embed_host.cpp', 'update_host.cpp', 'agg_host.cpp', 'emu_switch.cpp')]
DEPS = SRCS + [os.path.join(HERE, f) for f in ('cuda_runtime.h', 'cuda.h', 'emu_smem_ops.h', 'emu_tc_ops.h')] + [os.path.join(CSRC, f) for f in
                                                         ('deprel.cu', 'prune_csr.cu', 'batch.cu', 'pool3.cu', 'gemm_simt.cu', 'embed.cu', 'update.cu', 'aggregate.cu', 'tcgen05_util.cuh', 'gpt_common.cuh')]


def build():
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(['g++', '-std=c++17', '-O2', '-fPIC', '-shared', '-fno-extern-tls-init', '-Wno-psabi', '-x', 'c++', '-I', HERE] +
                          SRCS + ['-o', OUT])
    return OUT


if __name__ == '__main__':
    print(build())
