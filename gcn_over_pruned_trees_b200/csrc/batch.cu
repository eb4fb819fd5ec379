// K9 -- device-resident batch builder: the padded, length-sorted batch tensors of the reference's loader, assembled
// on the GPU from a token arena that is uploaded once.
//
// Replaces DataLoader.__getitem__ (/root/reference/data/loader.py:81-141, semeval_loader.py:75-119): per batch the
// reference zips Python lists, sorts them by length, draws word dropout token by token with np.random.random()
// (loader.py:181-188), pads seven LongTensors with get_long_tensor (loader.py:167-174, fill 0; 150 for the position
// fields, loader.py:125-126) and the training loop copies all of them to the device (trainer.py:60-71).  Here the
// pre-tokenised corpus lives in HBM as int32 arrays indexed by token; the host only knows the
// sentence lengths, so it can sort a batch and size it without touching the device, and one launch writes
//   words, pos, ner, deprel, head, subj_pos, obj_pos  int64 [B, T]     masks  bool [B, T] (words == 0)     rels int64 [B]
// No host-to-device copy on the step path.  Word dropout (train mode): a token that is not <UNK> becomes <UNK> with
// probability p, drawn from Philox keyed by (seed, stream, sentence id, token) -- same rule as loader.py:183-184,
// different random stream (the host path of data/loader.py reproduces numpy's stream when bit parity is wanted).
#include "gpt_common.cuh"

namespace {

constexpr int kFields = 7;     // words, pos, ner, deprel, head, subj_pos, obj_pos
constexpr int kBatchThreads = 128;

struct BatchParams {
    const int32_t* arena[kFields];   // token-indexed; [2] (ner) may be null: SemEval batches have no NER field
    long long* out[kFields];         // [B, T] each; out[2] null when arena[2] is
    const long long* offsets;        // [n_sentences + 1] first token of every sentence
    const int32_t* labels;           // [n_sentences]
    const int32_t* sel;              // [B] sentence ids of this batch, in output row order
    unsigned char* masks;            // [B, T]
    long long* rels;                 // [B]
    int B, T;
    unsigned thresh24;               // word dropout: drop when (24 random bits) < thresh24
    unsigned long long seed, stream;
};

__global__ void __launch_bounds__(kBatchThreads) build_batch_kernel(const BatchParams p) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * kBatchThreads + threadIdx.x;
    const int s = p.sel[b];
    const long long o = p.offsets[s];
    const int len = (int)(p.offsets[s + 1] - o);
    if (t == 0) p.rels[b] = p.labels[s];
    if (t >= p.T) return;
    const size_t at = (size_t)b * p.T + t;
    const bool in = t < len;
    long long w = in ? p.arena[0][o + t] : 0;                       // PAD_ID
    if (in && p.thresh24 > 0 && w != 1) {                           // UNK_ID = 1 stays what it is (loader.py:183)
        const Philox4 q = philox4x32((uint32_t)t, (uint32_t)s, (uint32_t)p.stream, (uint32_t)(p.stream >> 32),
                                     (uint32_t)p.seed, (uint32_t)(p.seed >> 32) ^ 0x574f5244u);
        if ((q.x >> 8) < p.thresh24) w = 1;
    }
    p.out[0][at] = w;
    p.masks[at] = in ? 0 : 1;                                       // torch.eq(words, 0), loader.py:108
#pragma unroll
    for (int f = 1; f < kFields; ++f) {
        if (p.arena[f] == nullptr) continue;
        const long long fill = f >= 5 ? 150 : 0;                    // position fields: loader.py:125-126
        p.out[f][at] = in ? (long long)p.arena[f][o + t] : fill;
    }
}

}  // namespace

extern "C" int gpt_build_batch(const int32_t* const* arena, const int64_t* offsets, const int32_t* labels,
                               const int32_t* sel, int B, int T, float word_dropout, uint64_t seed, uint64_t stream_id,
                               int64_t* const* out, uint8_t* masks, int64_t* rels, void* stream) {
    GPT_CHECK_ARG(arena && offsets && labels && sel && out && masks && rels && B >= 0 && T >= 1);
    GPT_CHECK_ARG(word_dropout >= 0.f && word_dropout < 1.f);
    if (B == 0) return GPT_OK;
    BatchParams p{};
    for (int f = 0; f < kFields; ++f) {
        p.arena[f] = arena[f];
        p.out[f] = reinterpret_cast<long long*>(out[f]);
        GPT_CHECK_ARG((arena[f] != nullptr) == (out[f] != nullptr));
        GPT_CHECK_ARG(f == 2 || arena[f] != nullptr);
    }
    p.offsets = reinterpret_cast<const long long*>(offsets);
    p.labels = labels; p.sel = sel; p.masks = masks; p.rels = reinterpret_cast<long long*>(rels);
    p.B = B; p.T = T;
    p.thresh24 = (unsigned)(word_dropout * 16777216.0f);
    p.seed = seed; p.stream = stream_id;
#ifdef GPT_HOST_EMULATION   // tests/emu: g++ has no <<<>>>
    gpt_launch(build_batch_kernel, dim3((T + kBatchThreads - 1) / kBatchThreads, B), dim3(kBatchThreads), 0,
               (cudaStream_t)stream, p);
#else
    build_batch_kernel<<<dim3((T + kBatchThreads - 1) / kBatchThreads, B), kBatchThreads, 0, (cudaStream_t)stream>>>(p);
#endif
    return gpt_launch_status();
}
