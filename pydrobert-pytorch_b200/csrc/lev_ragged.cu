// Ragged corpus -> padded batch_first token matrix (bulk-scoring front end).
//
// The reference builds its batches on the host, one utterance at a time
// (command_line.py:1110-1121: torch.tensor(transcript + [eos]) per utterance, then
// pad_sequence with padding_value=padding).  Here the corpus lives in HBM as two flat arrays
// (tokens, offsets); row u of the output is utterance sel[u] (or u), then eos, then padding.
// One thread per output element: the stores of a warp are one contiguous segment, the loads
// are contiguous within an utterance (and across utterances when sel is the identity).
#include "lev_common.cuh"

template <typename TT>
__global__ void __launch_bounds__(256)
lev_ragged_kernel(const TT* __restrict__ flat, const int64_t* __restrict__ off,
                  const int64_t* __restrict__ sel, int64_t n, int T, TT eos, TT pad,
                  TT* __restrict__ out) {
    const int64_t total = n * (int64_t)T;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool narrow = total < ((int64_t)1 << 31);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        int64_t u;
        if (narrow)
            u = (int64_t)((unsigned)idx / (unsigned)T);
        else
            u = idx / T;
        const int64_t j = idx - u * T;
        const int64_t s = sel != nullptr ? sel[u] : u;
        const int64_t b = off[s];
        const int64_t len = off[s + 1] - b;
        out[idx] = j < len ? flat[b + j] : (j == len ? eos : pad);
    }
}

template <typename TT>
static void lev_ragged_launch(const void* flat, const int64_t* off, const int64_t* sel, int64_t n,
                              int64_t T, int64_t eos, int64_t pad, void* out, cudaStream_t st) {
    const int64_t total = n * T;
    int64_t ctas = (total + 255) / 256;
    if (ctas > 148 * 64) ctas = 148 * 64;  // grid-stride beyond 8 resident CTAs x 8 rounds per SM
    lev_launch(lev_ragged_kernel<TT>, dim3((unsigned)ctas), dim3(256), 0, st, (const TT*)flat, off, sel,
               n, (int)T, (TT)eos, (TT)pad, (TT*)out);
}

extern "C" int b200lev_ragged_to_padded(const void* flat, int32_t elem_bytes, const int64_t* offsets,
                                        const int64_t* sel, int64_t n, int64_t T, int64_t eos,
                                        int64_t pad, void* out, void* stream) {
    if (n < 0 || T < 1 || T >= ((int64_t)1 << 31)) {
        lev_set_error("ragged_to_padded: n (%lld) must be >= 0 and T (%lld) in [1, 2^31)", (long long)n,
                      (long long)T);
        return B200LEV_ERR_ARG;
    }
    if (elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8) {
        lev_set_error("ragged_to_padded: elem_bytes (%d) must be 2, 4 or 8", (int)elem_bytes);
        return B200LEV_ERR_ARG;
    }
    if (offsets == nullptr || out == nullptr) {  // flat may be NULL: a corpus of empty utterances
        if (n == 0) return B200LEV_OK;
        lev_set_error("ragged_to_padded: NULL pointer");
        return B200LEV_ERR_ARG;
    }
    if (n == 0) return B200LEV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (elem_bytes == 2)
        lev_ragged_launch<int16_t>(flat, offsets, sel, n, T, eos, pad, out, st);
    else if (elem_bytes == 4)
        lev_ragged_launch<int32_t>(flat, offsets, sel, n, T, eos, pad, out, st);
    else
        lev_ragged_launch<int64_t>(flat, offsets, sel, n, T, eos, pad, out, st);
    return lev_check_cuda("lev_ragged_kernel");
}
