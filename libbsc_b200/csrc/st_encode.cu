// libbsc_b200/csrc/st_encode.cu -- Sort Transform of order k (k = 3..8) on the device.
//
// Replaces bsc_st_encode (libbsc/st/st.cpp:990-1012; CPU k=3..6 st.cpp:56-866; CUDA k=5..8
// st.cu:99-401).  Definition: P = STABLE sort of the n cyclic rotations by their
// first k bytes; L[j] = T[(P[j]-1) mod n]; index = j with P[j] = 0.
//
// One k-byte-key LSD radix sort.  The sort key of rotation i is the 64-bit big-endian word
//     k <= 7 :  T[i-1] | T[i] T[i+1] ... T[i+6]      (preceding byte rides in the top byte;
//               only the k context bytes, bits [(7-k)*8, 56), are sorted on)
//     k == 8 :  T[i] ... T[i+7], with T[i-1] carried as the value
// synthesised on the fly from a cyclically padded copy of the text by the first radix pass (the
// reference materialises 8n bytes of keys first, st.cu:99-147).  L is the top byte / the value of
// the sorted sequence.  Because the sort is stable and rotation 0 has the smallest position, it is
// the first element whose whole word equals rotation 0's word -> index by atomicMin (same
// argument as st.cu:149-163).
#include "radix_sort.cuh"
#include "stages.cuh"

namespace {

// Tw[0] = T[n-1]; Tw[1+i] = T[i]; Tw[1+n+j] = T[j mod n] (j < 16); then zeros.
__global__ void __launch_bounds__(256) st_build_wrap(const u8 *__restrict__ T, u8 *__restrict__ Tw, u32 n)
{
    u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) Tw[1 + i] = T[i];
    if (i < 16) Tw[1 + n + i] = T[i % n];
    if (i >= 16 && i < 48) Tw[1 + n + i] = 0;
    if (i == 0) Tw[0] = T[n - 1];
}

struct SrcST7 {
    const u8 *Tw;
    __device__ __forceinline__ u64 key(u32 i) const { return load_be64(Tw, i); }
    __device__ __forceinline__ u32 val(u32) const { return 0; }
};
struct SrcST8 {
    const u8 *Tw;
    __device__ __forceinline__ u64 key(u32 i) const { return load_be64(Tw, i + 1); }
    __device__ __forceinline__ u32 val(u32 i) const { return Tw[i]; }
};

template <bool K8>
__global__ void __launch_bounds__(256) st_finish(const u64 *__restrict__ keys, const u32 *__restrict__ vals, const u8 *__restrict__ Tw,
                                                 u8 *__restrict__ L, u32 n, u32 *index_out)
{
    u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    u64 lookup = load_be64(Tw, K8 ? 1u : 0u);
    u64 k = keys[j];
    L[j] = K8 ? (u8)vals[j] : (u8)(k >> 56);
    if (k == lookup) atomicMin(index_out, j);
}

}  // namespace

int stage_st_encode(Ctx *ctx, u8 *d_T, int n_, int k)
{
    if (n_ < 0) return LIBBSC_BAD_PARAMETER;
    if (k < 3 || k > 8) return LIBBSC_BAD_PARAMETER;
    if (n_ <= 1) return 0;
    const u32 n = (u32)n_;
    Arena &A = ctx->sort_arena();
    const size_t mark = A.mark();
    u8  *Tw   = A.get<u8>((size_t)n + 64);
    u64 *kb[2] = { A.get<u64>(n), A.get<u64>(n) };
    u32 *vb[2] = { nullptr, nullptr };
    if (k == 8) { vb[0] = A.get<u32>(n); vb[1] = A.get<u32>(n); }
    void *scratch = A.get<u8>(rs_scratch_bytes(n, RS_MAX_PASSES));

    LAUNCH(ctx, st_build_wrap, ceil_div(n, 256) , 256, 0, d_T, Tw, n);
    CUDA_TRY(cudaMemsetAsync(ctx->d_mail, 0xff, sizeof(u32), ctx->stream));
    int cur;
    if (k == 8) {
        cur = rs_sort<u64, true>(ctx, SrcST8{Tw}, kb, vb, n, make_passes(0, 64), scratch);
        LAUNCH(ctx, st_finish<true>, ceil_div(n, 256), 256, 0, kb[cur], vb[cur], Tw, d_T, n, ctx->d_mail);
    } else {
        cur = rs_sort<u64, false>(ctx, SrcST7{Tw}, kb, vb, n, make_passes((7 - k) * 8, 56), scratch);
        LAUNCH(ctx, st_finish<false>, ceil_div(n, 256), 256, 0, kb[cur], (const u32 *)nullptr, Tw, d_T, n, ctx->d_mail);
    }
    ctx->fetch_mail(1);
    A.release(mark);
    return (int)ctx->h_mail[0];
}
