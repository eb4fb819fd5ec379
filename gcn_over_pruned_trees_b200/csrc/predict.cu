// K11 -- the eval-side tail of GCNTrainer.predict (/root/reference/model/trainer.py:112-124) in one launch:
//
//   loss  = CrossEntropyLoss()(logits, labels)                     (mean over the batch, :118)
//   probs = F.softmax(logits, 1)                                    (:119)
//   predictions = np.argmax(logits, axis=1)                         (:120, first maximum wins)
//   unsort: rows re-ordered to the loader's original order          (:121-123, sorted(zip(orig_idx, ...)))
//
// The reference does this with three ATen launches, two device-to-host copies and a Python sort of B tuples.  Here
// one CTA walks the rows (a warp per row), writes probs / predictions straight to the row's ORIGINAL position
// (dest[b] = rank of orig_idx[b]) of one packed result buffer [ probs f32 B*C | predictions i32 B | loss f32 ], which
// the host reads back with a single copy.  The row losses are added in a fixed order (deterministic).
#include "gpt_common.cuh"

namespace {

constexpr int kPredThreads = 256;
constexpr int kPredWarps = kPredThreads / 32;

__global__ void __launch_bounds__(kPredThreads)
predict_tail_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, const int* __restrict__ dest,
                    int B, int C, float* __restrict__ probs, int* __restrict__ preds, float* __restrict__ loss) {
    GPT_PDL_ENTER();
    __shared__ float s_part[kPredWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float part = 0.f;                                   // this warp's rows, in row order
    for (int b = warp; b < B; b += kPredWarps) {
        const float* row = logits + (size_t)b * C;
        float mx = -INFINITY;
        int arg = 0x7fffffff;
        for (int c = lane; c < C; c += 32) {            // strict > keeps the first maximum of the lane's columns
            const float v = row[c];
            if (v > mx) { mx = v; arg = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(GPT_FULL_MASK, mx, o);
            const int oa = __shfl_xor_sync(GPT_FULL_MASK, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float sum = 0.f;
        for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
        sum = warp_sum_f(sum);
        const int d = dest != nullptr ? dest[b] : b;
        const float inv = 1.f / sum;
        for (int c = lane; c < C; c += 32) probs[(size_t)d * C + c] = expf(row[c] - mx) * inv;
        if (lane == 0) {
            preds[d] = arg;
            const long long y = labels[b];
            if (y >= 0 && y < C) part += logf(sum) + mx - row[y];        // -log softmax(row)[y]
        }
    }
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kPredWarps; ++w) t += s_part[w];
        *loss = B > 0 ? t / (float)B : 0.f;
    }
}

}  // namespace

extern "C" long long gpt_predict_result_bytes(int B, int C) {
    return (long long)B * C * 4 + (long long)B * 4 + 4;
}

extern "C" int gpt_predict_tail(const float* logits, const int64_t* labels, const int32_t* dest, int B, int C,
                                void* result, void* stream) {
    GPT_CHECK_ARG(logits && labels && result && B >= 0 && C >= 1);
    GPT_CHECK_ARG((reinterpret_cast<uintptr_t>(result) & 3) == 0);
    float* probs = reinterpret_cast<float*>(result);
    int* preds = reinterpret_cast<int*>(probs + (size_t)B * C);
    float* loss = reinterpret_cast<float*>(preds + B);
    gpt_launch(predict_tail_kernel, dim3(1), dim3(kPredThreads), 0, (cudaStream_t)stream, logits,
               reinterpret_cast<const long long*>(labels), dest, B, C, probs, preds, loss);
    return gpt_launch_status();
}
