// Dense linear layers / 1x1 convolutions on the 5th-generation tensor cores (sm_100a).
//
//   out[M, N] = epilogue( A[M, K] . B[N, K]^T + bias[N] )
//
// A = activations (tokens or NHWC pixels, K contiguous), B = the nn.Linear / 1x1-conv weight exactly as
// the checkpoint stores it ([out, in], K contiguous).  Two operand formats share the code (template KIND):
//   KIND 0: fp32 storage, TF32 tensor-core math (kind::tf32) — the precision PyTorch's own conv path uses on this GPU;
//   KIND 1: bf16 storage (activations + a bf16 copy of the weight), kind::f16 with bf16 operands.
// Accumulation is fp32 in TMEM either way; the output (and the residual read by the epilogue) is fp32 or bf16 (OutT),
// bias / LayerNorm parameters / statistics stay fp32.
//
// Replaces the cuBLAS GEMM + separate bias / ReLU / residual-add / LayerNorm kernels of the reference's
// projection and FFN layers (projects/SeqFormer/seqformer/models/deformable_transformer.py:155-169,
// ops/modules/ms_deform_attn.py:94-100,114) with ONE kernel per layer:
//   * warp 0   : TMA producer   (cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage mbarrier ring)
//   * warp 1   : MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M=128 x N=BLOCK_N x K=32B)
//   * warp 2   : TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps the MMAs of i+1)
//   * warps 4-7: epilogue       (tcgen05.ld: TMEM lane == output row, so each thread owns a whole row
//                                slice: bias, ReLU, and — when N == BLOCK_N == 256 — residual add +
//                                LayerNorm with thread-local statistics, no cross-thread reduction)
//   * persistent grid, static round-robin over (m, n) tiles, n fastest so neighbouring CTAs share A in L2.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace vnext {

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 2-CTA cluster helpers: one CTA's TMA load lands in the shared memory of every CTA in `mask` (same offset) and
// signals each one's mbarrier; tcgen05.commit can likewise arrive on the same barrier in every CTA of the mask.
__device__ __forceinline__ void tma_load_2d_mcast(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar,
                                                  uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
        "[%2], %5;" ::"r"(smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *tmap, const void *smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// shared -> global tile store (TMA); rows / columns beyond the tensor extent are clipped by the hardware
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tmap, const void *smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {   // arrives on `bar` when all prior MMAs of this thread retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <int KIND>   // 0 = tf32, 1 = f16/bf16
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets columns [c, c+32) of TMEM lane (warp%4)*32 + t
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
        "%25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
}
// fp16 with saturation to the largest finite value (an overflowing activation must not become inf)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(fminf(fmaxf(lo, -65504.f), 65504.f), fminf(fmaxf(hi, -65504.f), 65504.f));
    return *reinterpret_cast<const uint32_t *>(&h);
}

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B (8 rows x 128 B = 1024-B atoms):
//   start address >> 4 | LBO (ignored for swizzled K-major) = 1 | SBO = 1024 B between 8-row groups | version 1 | layout 2
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// c_format F32 (1) @4 | a_format @7 | b_format @10 | K-major A, B (0) @15,16 | N>>3 @17 | M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int M, int N) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
enum { EPI_RELU = 1, EPI_RESIDUAL = 2, EPI_LAYERNORM = 4, EPI_GELU = 8 };

struct GemmParams {
    int M, N, K;
    const float *bias;       // [N] or null
    const float *bias2d;     // [period, ldb2] or null: added as bias2d[row % period][col] (geometry-dependent bias)
    int period, ldb2;
    int pdl;                 // launched with programmatic stream serialization (see the kernel prologue)
    int f16_in, f16_out;     // the 16-bit operands / outputs are IEEE fp16 instead of bf16 (same kernels, other formats)
    // MSDA front-end in the epilogue (N == 384 = 8 heads x (4 levels x 4 points x 2 offsets | 4 x 4 logits)): the offsets
    // columns become sampling locations ref + off / (W_l, H_l), the logits columns the softmax over each head's 16 — what
    // MSDeformAttn.forward does between its projection and the sampler (ops/modules/ms_deform_attn.py:98-108)
    const float *msda_ref;   // [period, 4 levels, 2] reference points (row % period), or null
    float msda_wh[8];        // (1/W_0, 1/H_0, ..., 1/W_3, 1/H_3)
    int b2_tma;              // the 2-D bias tile comes through shared memory (four 8-row TMA boxes, rows taken mod period)
    const unsigned char *row_mask;   // [period] or null: rows with mask != 0 get columns [0, mask_cols) zeroed
    int mask_cols;
    const float *residual;   // [M, ldr] (EPI_RESIDUAL)
    const float *gamma;      // [N]      (EPI_LAYERNORM)
    const float *beta;       // [N]
    float *out;              // [M, ldo]
    int ldo, ldr;
    int flags;
    int vec_ok;              // out / residual / bias rows allow 16-byte vector access
    int tma_epi;             // epilogue tiles go through shared memory + TMA (coalesced); else direct row stores
    float eps;
    int num_m_tiles, num_n_tiles;
    // implicit-GEMM 3x3 convolution (stride 1, zero padding 1) over NHWC: the A tile of an 8 x 16 pixel patch and filter
    // tap (ky, kx) is ONE 4-D TMA box at (channel block, x0 + kx - 1, y0 + ky - 1, frame); out-of-image pixels are
    // zero-filled by the TMA unit, which is exactly the convolution's padding
    int conv, conv_cin_blocks, conv_tiles_x, conv_tiles_y, conv_stride;
    // cluster == 2: CTA pairs work on two vertically adjacent m-tiles of the same n-tile; each CTA loads half of the
    // weight (B) tile and multicasts it to both, halving the dominant L2 -> shared-memory operand traffic
    int cluster;
};

constexpr int kGemmThreads = 384;      // warps 0-3: TMA / MMA / TMEM-alloc / idle; warps 4-11: epilogue
constexpr int kEpiWarps = 8;           // two warps per TMEM lane quarter, interleaving the 32-column chunks
constexpr int BLOCK_M = 128;

// x[j] = acc + bias + bias2d + residual for 32 consecutive columns starting at global column `col0` of `row`
// the parts of the epilogue sum that still come from global memory: row-periodic 2-D bias, and the residual when
// the TMA path is unavailable (unaligned rows)
__device__ __forceinline__ void epilogue_add_global(const GemmParams &p, int row, bool row_ok, int col0, uint32_t (&v)[32]) {
    if (!row_ok) return;
    const float *b2 = (p.bias2d && !p.b2_tma) ? p.bias2d + (size_t)(row % p.period) * p.ldb2 + col0 : nullptr;
    const float *rr = ((p.flags & EPI_RESIDUAL) && !p.tma_epi) ? p.residual + (size_t)row * p.ldr + col0 : nullptr;
    const bool vec = p.vec_ok && (col0 + 32 <= p.N);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float x[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])};
        if (vec) {
            if (b2) { const float4 b = *reinterpret_cast<const float4 *>(b2 + j); x[0] += b.x; x[1] += b.y; x[2] += b.z; x[3] += b.w; }
            if (rr) { const float4 r = *reinterpret_cast<const float4 *>(rr + j); x[0] += r.x; x[1] += r.y; x[2] += r.z; x[3] += r.w; }
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (col0 + j + t < p.N) {
                    if (b2) x[t] += b2[j + t];
                    if (rr) x[t] += rr[j + t];
                }
        }
        v[j] = __float_as_uint(x[0]); v[j + 1] = __float_as_uint(x[1]); v[j + 2] = __float_as_uint(x[2]); v[j + 3] = __float_as_uint(x[3]);
    }
}

template <int BLOCK_N, int STAGES, int KIND, bool OUT16>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
               const GemmParams p) {
    constexpr int BLOCK_K = KIND == 0 ? 32 : 64;   // elements per 128-byte swizzle row (fp32 / bf16)
    constexpr int UMMA_K = KIND == 0 ? 8 : 16;     // 32 bytes per MMA either way
    constexpr int A_BYTES = BLOCK_M * 128, B_BYTES = BLOCK_N * 128;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int ACC_STAGES = 2;
    constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;          // 256 or 512: power of two >= 32
    constexpr uint32_t IDESC = make_idesc(KIND == 0 ? 2 : 1, BLOCK_M, BLOCK_N);   // a/b format: 2 = TF32, 1 = BF16
    constexpr uint32_t EPI_TILE_BYTES = OUT16 ? 2048u : 4096u;   // 32 rows x 32 columns of OutT
    static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full[ACC_STAGES], tmem_empty[ACC_STAGES];
    __shared__ __align__(8) uint64_t res_bar[kEpiWarps];       // per epilogue warp: its residual staging tile
    __shared__ __align__(16) float s_bias[3][256];             // bias / gamma / beta of the current n-tile
    __shared__ float2 s_lnstat[kEpiWarps][32];                 // LayerNorm partial (sum, sumsq) exchange between warp halves
    __shared__ uint32_t tmem_base_smem;
    // epilogue staging after the operand ring: per warp 2 output + 2 residual tiles of 32 rows x 128 B (swizzled)
    unsigned char *const epi_smem = smem + STAGES * STAGE_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k_blocks = p.conv == 2 ? 7 : (p.conv ? 9 * p.conv_cin_blocks : (p.K + BLOCK_K - 1) / BLOCK_K);
    // work distribution: a work item is one output tile (cluster 1) or a vertical pair of tiles (cluster 2)
    const int csize = p.cluster;
    const uint32_t crank = csize > 1 ? cluster_ctarank() : 0u;
    const int first_work = blockIdx.x / csize, work_stride = gridDim.x / csize;
    const int num_work = ((p.num_m_tiles + csize - 1) / csize) * p.num_n_tiles;
    auto work_tile = [&](int w) { return ((w / p.num_n_tiles) * csize + (int)crank) * p.num_n_tiles + w % p.num_n_tiles; };

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], csize);
        }
#pragma unroll
        for (int s = 0; s < ACC_STAGES; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], kEpiWarps);
        }
        for (int w = 0; w < kEpiWarps; ++w) mbar_init(&res_bar[w], 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();          // peers' barriers are initialised before anything remote touches them
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (p.pdl) {
        // programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) ran while the
        // previous kernel of the stream was still finishing; its results are needed from here on.  Our own dependents may
        // start their prologue right away — they wait for this grid's completion at the same point.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int stage = 0, phase = 0;
        for (int work = first_work; work < num_work; work += work_stride) {
            const int tile = work_tile(work);
            const int m_idx = (tile / p.num_n_tiles) * BLOCK_M, n_idx = (tile % p.num_n_tiles) * BLOCK_N;
            // every CTA (pair) walks K from a different starting block: all SMs want the same weight tile at about
            // the same time, and starting them at different k spreads those requests over the L2 slices
            const int k_rot = (int)((unsigned)first_work % (unsigned)num_k_blocks);
            for (int kb0 = 0; kb0 < num_k_blocks; ++kb0) {
                int kb = kb0 + k_rot;
                if (kb >= num_k_blocks) kb -= num_k_blocks;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char *sa = smem + stage * STAGE_BYTES, *sb = sa + A_BYTES;
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                if (p.conv == 2) {
                    // 7x7 stride-2 stem on a zero-bordered 4-channel image: k-block = filter row ky; the A tile is 128
                    // consecutive output pixels x (8 input pixels x 4 channels = 32 floats, the 8th pixel meets zero weights);
                    // neighbouring output pixels start 2 input pixels = 32 B apart: an OVERLAPPING tensor-map view (see
                    // stem_conv_device) delivers the im2col rows without an im2col buffer
                    const int mt = tile / p.num_n_tiles;
                    const int tx = mt % p.conv_tiles_x, ty = (mt / p.conv_tiles_x) % p.conv_tiles_y;
                    tma_load_4d(sa, &tmap_a, 0, tx * 128, 2 * ty + kb, mt / (p.conv_tiles_x * p.conv_tiles_y), &full_bar[stage]);
                } else if (p.conv) {
                    const int mt = tile / p.num_n_tiles;
                    const int tx = mt % p.conv_tiles_x, ty = (mt / p.conv_tiles_x) % p.conv_tiles_y;
                    const int fr = mt / (p.conv_tiles_x * p.conv_tiles_y);
                    const int tap = kb / p.conv_cin_blocks, cb = kb - tap * p.conv_cin_blocks;
                    // stride s: the box starts at input pixel (s*x0 + kx - 1, s*y0 + ky - 1) and the tensor map steps by s
                    tma_load_4d(sa, &tmap_a, cb * BLOCK_K, tx * 16 * p.conv_stride + tap % 3 - 1,
                                ty * 8 * p.conv_stride + tap / 3 - 1, fr, &full_bar[stage]);
                } else {
                    tma_load_2d(sa, &tmap_a, kb * BLOCK_K, m_idx, &full_bar[stage]);
                }
                if (csize > 1)      // my half of the weight tile, delivered to both CTAs of the pair
                    tma_load_2d_mcast(sb + crank * (B_BYTES / 2), &tmap_b, kb * BLOCK_K, n_idx + (int)crank * (BLOCK_N / 2),
                                      &full_bar[stage], (uint16_t)3);
                else
                    tma_load_2d(sb, &tmap_b, kb * BLOCK_K, n_idx, &full_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (KIND == 1 && p.f16_in) ? make_idesc(0, BLOCK_M, BLOCK_N) : IDESC;   // kind::f16 a/b format 0 = F16
        int stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int work = first_work; work < num_work; work += work_stride) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + A_BYTES;
                const uint64_t da = make_smem_desc_sw128(sa), db = make_smem_desc_sw128(sb);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k)   // advance 32 B inside the 128-B swizzle row: +2 in 16-B units
                    umma<KIND>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                if (csize > 1) umma_commit_mcast(&empty_bar[stage], (uint16_t)3);   // frees the slot in BOTH CTAs
                else umma_commit(&empty_bar[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tmem_full[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (8 warps) =====================
        const int ewi = warp - 4;                                   // 0..7
        const int ew = ewi & 3;                                     // TMEM lane quarter this warp may touch (warp % 4)
        const int half = ewi >> 2;                                  // which interleaved half of the column chunks
        int acc = 0, acc_phase = 0;
        const bool relu = (p.flags & EPI_RELU) != 0;
        const bool has_res = (p.flags & EPI_RESIDUAL) != 0;
        const bool tma_epi = p.tma_epi != 0;
        const bool b2_tma = p.b2_tma != 0;                          // (never together with a residual)
        const bool addend = has_res || b2_tma;                      // a 32 x 32 tile added to the accumulator, staged by TMA

        const bool ln = (p.flags & EPI_LAYERNORM) != 0;
        // staging: two 4-KB tiles (32 rows x 128 B, SWIZZLE_128B) per warp.  With a residual: [0] = out, [1] = residual;
        // without: both are output buffers (double-buffered TMA stores).
        unsigned char *const buf0 = epi_smem + ewi * 8192;
        unsigned char *const buf1 = buf0 + 4096;
        // one addend tile -> buf1: the residual's rows row0..row0+31, or the periodic bias rows (row0 + r) % period, which
        // are four 8-row boxes because the period need not be a multiple of 32 (it is a multiple of 8, checked on the host)
        auto load_addend = [&](int col, int row0_) {
            mbar_expect_tx(&res_bar[ewi], b2_tma ? 4096u : EPI_TILE_BYTES);
            if (b2_tma) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    tma_load_2d(buf1 + q * 1024, &tmap_res, col, (row0_ + 8 * q) % p.period, &res_bar[ewi]);
            } else {
                tma_load_2d(buf1, &tmap_res, col, row0_, &res_bar[ewi]);
            }
        };
        const int swz = lane & 7;                                   // 16-B chunk j of row r lives at j ^ (r & 7)
        const int swz64 = (lane >> 1) & 3;                          // same for the 64-B rows of a bf16 tile
        uint32_t stores = 0, res_loads = 0;
        int bias_n_idx = -1;
        if (tma_epi && lane == 0) {
            prefetch_tmap(&tmap_out);
            if (addend) prefetch_tmap(&tmap_res);
        }
        for (int work = first_work; work < num_work; work += work_stride) {
            const int tile = work_tile(work);
            const int m_idx = (tile / p.num_n_tiles) * BLOCK_M, n_idx = (tile % p.num_n_tiles) * BLOCK_N;
            const int row0 = m_idx + ew * 32;
            const int row = row0 + lane;
            const bool row_ok = row < p.M;
            const int n_chunks = min(BLOCK_N, p.N - n_idx + 31) / 32;   // 32-column chunks that hold real columns
            if (n_idx != bias_n_idx) {                                  // (re)load the per-column vectors of this n-tile
                asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone is done with the previous vectors
                const int t = threadIdx.x - 128;                        // 0..255
                const int col = n_idx + t;
                if (t < BLOCK_N) {
                    s_bias[0][t] = (p.bias && col < p.N) ? p.bias[col] : 0.f;
                    if (ln) {
                        s_bias[1][t] = col < p.N ? p.gamma[col] : 0.f;
                        s_bias[2][t] = col < p.N ? p.beta[col] : 0.f;
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                bias_n_idx = n_idx;
            }
            if (tma_epi && addend && half < n_chunks) {                 // first addend tile: in flight while we wait
                if (lane == 0) load_addend(n_idx + half * 32, row0);
            }
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BLOCK_N;
            float *orow = p.out + (size_t)row * p.ldo;
            const bool masked = p.row_mask && row_ok && p.row_mask[row % p.period] != 0;
            float mean = 0.f, rstd = 1.f;
            // LayerNorm: pass 0 parks x = acc + bias (+ residual) in TMEM and gathers thread-local statistics
            // (TMEM lane == output row; the two warps of a lane quarter swap partial sums); pass 1 normalises and stores.
            for (int pass = ln ? 0 : 1; pass < 2; ++pass) {
                float sum = 0.f, sumsq = 0.f;
#pragma unroll 1
                for (int c = half; c < n_chunks; c += 2) {
                    const int c0 = c * 32, col0 = n_idx + c0;
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    if ((pass == 0 || !ln) && !(p.flags & 0x200)) {
                        const unsigned char *rrow = nullptr;
                        if (tma_epi && addend) {
                            mbar_wait(&res_bar[ewi], res_loads & 1);
                            ++res_loads;
                            rrow = buf1 + lane * ((OUT16 && !b2_tma) ? 64 : 128);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {                    // bias from shared memory (broadcast reads)
                            const float4 b = *reinterpret_cast<const float4 *>(&s_bias[0][c0 + 4 * j]);
                            float4 x = make_float4(__uint_as_float(v[4 * j]) + b.x, __uint_as_float(v[4 * j + 1]) + b.y,
                                                   __uint_as_float(v[4 * j + 2]) + b.z, __uint_as_float(v[4 * j + 3]) + b.w);
                            if (rrow) {
                                float4 r;
                                if (OUT16 && !b2_tma)     // 64-B rows, SWIZZLE_64B: 16-B chunk c of row r lives at c ^ ((r >> 1) & 3)
                                    r = load4(reinterpret_cast<const __nv_bfloat16 *>(rrow + (((j >> 1) ^ swz64) << 4) + (j & 1) * 8));
                                else
                                    r = *reinterpret_cast<const float4 *>(rrow + ((j ^ swz) << 4));
                                x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
                            }
                            v[4 * j] = __float_as_uint(x.x); v[4 * j + 1] = __float_as_uint(x.y);
                            v[4 * j + 2] = __float_as_uint(x.z); v[4 * j + 3] = __float_as_uint(x.w);
                        }
                        if (rrow) {
                            __syncwarp();                                // tile consumed by all lanes -> refill it
                            if (lane == 0 && c + 2 < n_chunks) load_addend(col0 + 64, row0);
                        }
                        if ((p.bias2d && !b2_tma) || (has_res && !rrow)) epilogue_add_global(p, row, row_ok, col0, v);
                        if (p.msda_ref && row_ok) {
                            if (col0 < 256) {        // column = ((head*4 + level)*4 + point)*2 + xy  ->  level = (j >> 3) & 3
                                const float4 ra = *reinterpret_cast<const float4 *>(p.msda_ref + (size_t)(row % p.period) * 8);
                                const float4 rb = *reinterpret_cast<const float4 *>(p.msda_ref + (size_t)(row % p.period) * 8 + 4);
                                const float rf[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int k = ((j >> 3) & 3) * 2 + (j & 1);
                                    v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), p.msda_wh[k], rf[k]));
                                }
                            } else {                 // two heads' 16 logits each: softmax in registers
#pragma unroll
                                for (int hh = 0; hh < 2; ++hh) {
                                    float mx = -INFINITY;
#pragma unroll
                                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[hh * 16 + j]));
                                    float sm = 0.f;
#pragma unroll
                                    for (int j = 0; j < 16; ++j) {
                                        const float e = __expf(__uint_as_float(v[hh * 16 + j]) - mx);
                                        sm += e;
                                        v[hh * 16 + j] = __float_as_uint(e);
                                    }
#pragma unroll
                                    const float inv = __frcp_rn(sm);
#pragma unroll
                                    for (int j = 0; j < 16; ++j) v[hh * 16 + j] = __float_as_uint(__uint_as_float(v[hh * 16 + j]) * inv);
                                }
                            }
                        }
                    }
                    if (pass == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = __uint_as_float(v[j]);
                            sum += x;
                            sumsq = fmaf(x, x, sumsq);
                        }
                        tmem_st32(taddr + c0, v);
                        continue;
                    }
                    if (ln) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 g = *reinterpret_cast<const float4 *>(&s_bias[1][c0 + j]);
                            const float4 be = *reinterpret_cast<const float4 *>(&s_bias[2][c0 + j]);
                            v[j] = __float_as_uint((__uint_as_float(v[j]) - mean) * rstd * g.x + be.x);
                            v[j + 1] = __float_as_uint((__uint_as_float(v[j + 1]) - mean) * rstd * g.y + be.y);
                            v[j + 2] = __float_as_uint((__uint_as_float(v[j + 2]) - mean) * rstd * g.z + be.z);
                            v[j + 3] = __float_as_uint((__uint_as_float(v[j + 3]) - mean) * rstd * g.w + be.w);
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
                    }
                    if (p.flags & EPI_GELU) {                            // exact (erf) GELU, nn.GELU's default
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = __uint_as_float(v[j]);
                            v[j] = __float_as_uint(0.5f * x * (1.f + erff(x * 0.70710678118654752f)));
                        }
                    }
                    if (masked) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.mask_cols) v[j] = 0u;
                    }
                    if (p.flags & 0x100) continue;                       // debug: no output at all
                    if (tma_epi) {
                        unsigned char *ob = (addend || !(stores & 1)) ? buf0 : buf1;
                        const int pending_ok = addend ? 0 : 1;           // groups that may still be reading shared memory
                        if (stores > (uint32_t)pending_ok) {
                            if (lane == 0) {
                                if (pending_ok) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
                            }
                            __syncwarp();
                        }
                        if (OUT16 && p.f16_out) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 pk;
                                pk.x = pack_f16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
                                pk.y = pack_f16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                                pk.z = pack_f16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                                pk.w = pack_f16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                                *reinterpret_cast<uint4 *>(ob + lane * 64 + ((j ^ swz64) << 4)) = pk;
                            }
                        } else if (OUT16) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 pk;
                                pk.x = pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
                                pk.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                                pk.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                                pk.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                                *reinterpret_cast<uint4 *>(ob + lane * 64 + ((j ^ swz64) << 4)) = pk;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                *reinterpret_cast<float4 *>(ob + lane * 128 + ((j ^ swz) << 4)) =
                                    make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            if (p.conv == 2) {   // this warp's 32 tile rows = 32 consecutive pixels of one output row
                                const int mt = tile / p.num_n_tiles;
                                const int tx = mt % p.conv_tiles_x, ty = (mt / p.conv_tiles_x) % p.conv_tiles_y;
                                tma_store_4d(&tmap_out, ob, col0, tx * 128 + ew * 32, ty, mt / (p.conv_tiles_x * p.conv_tiles_y));
                            } else if (p.conv) {     // this warp's 32 tile rows = 2 image rows x 16 pixels of the 8 x 16 patch
                                const int mt = tile / p.num_n_tiles;
                                const int tx = mt % p.conv_tiles_x, ty = (mt / p.conv_tiles_x) % p.conv_tiles_y;
                                tma_store_4d(&tmap_out, ob, col0, tx * 16, ty * 8 + 2 * ew, mt / (p.conv_tiles_x * p.conv_tiles_y));
                            } else {
                                tma_store_2d(&tmap_out, ob, col0, row0);
                            }
                            tma_store_commit();
                        }
                        ++stores;
                    } else if (row_ok && !OUT16) {       // (bf16 output always takes the TMA path; checked on the host)
                        if (p.vec_ok && col0 + 32 <= p.N) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4 *>(orow + col0 + j) =
                                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.N) orow[col0 + j] = __uint_as_float(v[j]);
                        }
                    }
                }
                if (pass == 0) {
                    tmem_st_wait();
                    s_lnstat[ewi][lane] = make_float2(sum, sumsq);      // swap partials with the other half's warp
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + ew) : "memory");
                    const float2 o = s_lnstat[ewi ^ 4][lane];
                    sum += o.x;
                    sumsq += o.y;
                    mean = sum * (1.f / BLOCK_N);
                    const float var = fmaxf(sumsq * (1.f / BLOCK_N) - mean * mean, 0.f);
                    rstd = rsqrtf(var + p.eps);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (tma_epi && lane == 0) tma_store_wait_all();             // stores must land before the CTA exits
    }
    tcgen05_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all();          // nobody leaves while a peer may still multicast into it / arrive on its barriers
    if (warp == 2) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// 2-D row-major [rows, cols] tensor of fp32 (dt 0) or bf16 (dt 2) with row pitch ld (elements); box = [box_rows,
// box_cols].  Operand tiles: 128-byte rows (32 fp32 / 64 bf16 columns), SWIZZLE_128B.  Epilogue tiles: 32 columns,
// SWIZZLE_128B for fp32 (128-B rows) / SWIZZLE_64B for bf16 (64-B rows).
static int make_tmap_2d(CUtensorMap *tm, int dt, const void *ptr, int rows, int cols, int ld, int box_rows, int box_cols) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VNEXT_ECUDA;
    }
    const int esz = dt == VNEXT_BF16 || dt == VNEXT_F16 ? 2 : 4;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(tm, dt == VNEXT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : dt == VNEXT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                             : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_cols * esz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d dtype %d", (int)r, rows, cols, ld, dt);
        return VNEXT_ECUDA;
    }
    return VNEXT_OK;
}

// 4-D NHWC tensor [T, H, W, C]; box = [1, box_h, box_w, box_c channels]
static int make_tmap_nhwc(CUtensorMap *tm, int dt, const void *ptr, int T, int H, int W, int C, int box_h, int box_w,
                          int box_c, int estride = 1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VNEXT_ECUDA;
    }
    const cuuint64_t esz = dt == VNEXT_BF16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T};
    cuuint64_t strides[3] = {(cuuint64_t)C * esz, (cuuint64_t)W * C * esz, (cuuint64_t)H * W * C * esz};
    // element strides (1, s, s, 1): a box of box_w*s x box_h*s input pixels delivers box_w x box_h of them
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), 1u};
    cuuint32_t estr[4] = {1u, (cuuint32_t)estride, (cuuint32_t)estride, 1u};
    CUresult r = enc(tm, dt == VNEXT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     box_c * esz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (NHWC) failed (%d) for [%d,%d,%d,%d] dtype %d", (int)r, T, H, W, C, dt);
        return VNEXT_ECUDA;
    }
    return VNEXT_OK;
}

template <int BLOCK_N, int STAGES, int KIND, bool OUT16>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &to, const CUtensorMap &tr,
                       GemmParams p, cudaStream_t stream) {
    auto kern = gemm_tc_kernel<BLOCK_N, STAGES, KIND, OUT16>;
    constexpr size_t smem = (size_t)STAGES * (BLOCK_M * 128 + BLOCK_N * 128) + kEpiWarps * 8192 + 1024;
    static bool configured = false;
    if (!configured) {
        VNEXT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    if (!p.conv) p.num_m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    p.num_n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    int grid = tiles < num_sms() ? tiles : num_sms();
    if (p.cluster == 2) {
        const int pairs = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
        grid = 2 * (pairs < num_sms() / 2 ? pairs : num_sms() / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        VNEXT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tr, p));
        count_launch();
        return VNEXT_OK;
    }
    static const bool want_pdl = getenv("VNEXT_GEMM_PDL") ? atoi(getenv("VNEXT_GEMM_PDL")) != 0 : false;
    if (want_pdl) {
        p.pdl = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        VNEXT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tr, p));
        count_launch();
        return VNEXT_OK;
    }
    kern<<<grid, kGemmThreads, smem, stream>>>(ta, tb, to, tr, p);
    count_launch();
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}

// operand format x output format -> instantiation
static int launch_any(bool wide, int in_dt, int out_dt, const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &to,
                      const CUtensorMap &tr, const GemmParams &p, cudaStream_t stream) {
    if (in_dt == VNEXT_F16) in_dt = VNEXT_BF16;        // same instantiations: the format is a run-time field (p.f16_in / p.f16_out)
    if (out_dt == VNEXT_F16) out_dt = VNEXT_BF16;
    if (in_dt == VNEXT_F32 && out_dt == VNEXT_F32)
        return wide ? launch_gemm<256, 3, 0, false>(ta, tb, to, tr, p, stream) : launch_gemm<128, 4, 0, false>(ta, tb, to, tr, p, stream);
    if (in_dt == VNEXT_F32 && out_dt == VNEXT_BF16)
        return wide ? launch_gemm<256, 3, 0, true>(ta, tb, to, tr, p, stream) : launch_gemm<128, 4, 0, true>(ta, tb, to, tr, p, stream);
    if (in_dt == VNEXT_BF16 && out_dt == VNEXT_F32)
        return wide ? launch_gemm<256, 3, 1, false>(ta, tb, to, tr, p, stream) : launch_gemm<128, 4, 1, false>(ta, tb, to, tr, p, stream);
    if (in_dt == VNEXT_BF16 && out_dt == VNEXT_BF16)
        return wide ? launch_gemm<256, 3, 1, true>(ta, tb, to, tr, p, stream) : launch_gemm<128, 4, 1, true>(ta, tb, to, tr, p, stream);
    set_error("vnext_linear: unsupported dtypes (in %d, out %d)", in_dt, out_dt);
    return VNEXT_EINVAL;
}

struct LinearArgs {
    int in_dt, out_dt;       // VNEXT_F32 / VNEXT_BF16: a + weight, out + residual
    const void *a; int lda;
    const void *w;
    const float *bias;
    const float *bias2d; int period, ldb2;
    const unsigned char *row_mask; int mask_cols;
    const void *residual; int ldr;
    const float *gamma, *beta; float eps;
    void *out; int ldo;
    int M, N, K, flags;
    const float *msda_ref = nullptr;       // see GemmParams::msda_ref
    const float *msda_wh = nullptr;        // host pointer to 8 floats
};

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool is16(int dt) { return dt == VNEXT_BF16 || dt == VNEXT_F16; }     // 2-byte storage types

int linear_tc_device(const LinearArgs &x, cudaStream_t stream) {
    if (x.M == 0 || x.N == 0) return VNEXT_OK;
    const int oe = is16(x.out_dt) ? 8 : 4;                   // elements per 16 bytes of out / residual
    GemmParams p;
    p.f16_in = x.in_dt == VNEXT_F16;
    p.f16_out = x.out_dt == VNEXT_F16;
    if (p.f16_out && x.residual) {
        set_error("vnext_linear: an fp16 output cannot carry a residual (the residual is read in the output's type; use fp32 / bf16)");
        return VNEXT_EINVAL;
    }
    p.M = x.M; p.N = x.N; p.K = x.K; p.bias = x.bias; p.bias2d = x.bias2d; p.period = x.period > 0 ? x.period : 1;
    p.ldb2 = x.ldb2; p.row_mask = x.row_mask; p.mask_cols = x.mask_cols; p.residual = static_cast<const float *>(x.residual);
    p.gamma = x.gamma; p.beta = x.beta; p.out = static_cast<float *>(x.out); p.ldo = x.ldo; p.ldr = x.ldr; p.flags = x.flags;
    p.eps = x.eps;
    p.num_m_tiles = p.num_n_tiles = 0;
    p.conv = p.conv_cin_blocks = p.conv_tiles_x = p.conv_tiles_y = 0;
    p.conv_stride = 1;
    p.cluster = 1;
    p.pdl = 0;
    p.msda_ref = x.msda_ref;
    for (int i = 0; i < 8; ++i) p.msda_wh[i] = x.msda_wh ? 1.f / x.msda_wh[i] : 1.f;
    p.b2_tma = 0;
    p.vec_ok = al16(x.out) && x.ldo % oe == 0 && (!x.bias || al16(x.bias)) &&
               (!x.bias2d || (al16(x.bias2d) && x.ldb2 % 4 == 0)) && (!x.residual || (al16(x.residual) && x.ldr % oe == 0));
    // 256-wide tiles only when N is a multiple of 256.  Measured (720p encoder offsets|logits projection, N = 384):
    // 256 + half-empty 256 tiles take 103 us per call against 84 us for three 128-wide tiles, so partial wide tiles lose.
    const bool wide = (x.N % 256 == 0);
    if (x.flags & EPI_LAYERNORM) {
        if (x.N != 256 || !p.vec_ok || !x.gamma || !x.beta || !al16(x.gamma) || !al16(x.beta)) {
            set_error("vnext_linear: the fused LayerNorm epilogue needs N == 256 (got %d) and 16-B aligned operands", x.N);
            return VNEXT_EINVAL;
        }
    }
    const int kbox = is16(x.in_dt) ? 64 : 32;
    CUtensorMap ta, tb, to, tr;
    int rc = make_tmap_2d(&ta, x.in_dt, x.a, x.M, x.K, x.lda, BLOCK_M, kbox);
    if (rc) return rc;
    {   // pair CTAs (weight-tile multicast) whenever there are enough m-tiles to keep every SM busy in pairs
        static const char *env = getenv("VNEXT_GEMM_CLUSTER");
        const int want = env ? atoi(env) : 1;      // measured: no gain (these GEMMs are HBM-bound, not operand-bound) -> off
        const int m_tiles = (x.M + BLOCK_M - 1) / BLOCK_M;
        if (want == 2 && m_tiles * ((x.N + (wide ? 255 : 127)) / (wide ? 256 : 128)) >= 2 * num_sms()) p.cluster = 2;
    }
    rc = make_tmap_2d(&tb, x.in_dt, x.w, x.N, x.K, x.K, (wide ? 256 : 128) / p.cluster, kbox);
    if (rc) return rc;
    // coalesced epilogue through shared memory + TMA needs 16-byte aligned rows of out (and residual)
    p.tma_epi = al16(x.out) && x.ldo % oe == 0 && (!x.residual || (al16(x.residual) && x.ldr % oe == 0));
    if ((x.flags & 0x400) && x.out_dt == VNEXT_F32) p.tma_epi = 0;   // debug: force the direct-store epilogue
    if (is16(x.out_dt) && !p.tma_epi) {
        set_error("vnext_linear: bf16 / fp16 output / residual rows must be 16-byte aligned (ldo %d, ldr %d)", x.ldo, x.ldr);
        return VNEXT_EINVAL;
    }
    to = ta;
    tr = ta;
    if (p.tma_epi) {
        rc = make_tmap_2d(&to, x.out_dt, x.out, x.M, x.N, x.ldo, 32, 32);
        if (rc) return rc;
        if (x.residual) {
            rc = make_tmap_2d(&tr, x.out_dt, x.residual, x.M, x.N, x.ldr, 32, 32);
            if (rc) return rc;
        } else if (x.bias2d && p.period % 8 == 0 && al16(x.bias2d) && x.ldb2 % 4 == 0 && x.N % 4 == 0) {
            // row-periodic 2-D bias as a TMA-staged addend (8-row boxes: a 32-row tile may wrap around the period)
            rc = make_tmap_2d(&tr, VNEXT_F32, x.bias2d, p.period, x.N, x.ldb2, 8, 32);
            if (rc) return rc;
            p.b2_tma = 1;
        }
    }
    return launch_any(wide, x.in_dt, x.out_dt, ta, tb, to, tr, p, stream);
}

int conv3x3_tc_device(int in_dt, int out_dt, const void *x, const void *w, const float *bias, void *out, int T, int Hin, int Win,
                      int Cin, int Cout, int stride, int relu, cudaStream_t stream) {
    const int kbox = in_dt == VNEXT_BF16 ? 64 : 32;
    const int H = (Hin - 1) / stride + 1, W = (Win - 1) / stride + 1;       // output size (kernel 3, padding 1)
    GemmParams p;
    p.conv_stride = stride;
    p.M = T * H * W; p.N = Cout; p.K = 9 * Cin; p.bias = bias; p.bias2d = nullptr; p.period = 1; p.ldb2 = 0;
    p.b2_tma = 0;
    p.pdl = 0;
    p.f16_in = p.f16_out = 0;
    p.msda_ref = nullptr;
    p.row_mask = nullptr; p.mask_cols = 0; p.residual = nullptr; p.gamma = p.beta = nullptr; p.out = static_cast<float *>(out);
    p.ldo = Cout; p.ldr = 0; p.flags = relu ? EPI_RELU : 0; p.eps = 0.f; p.vec_ok = 1; p.tma_epi = 1;
    p.conv = 1; p.conv_cin_blocks = Cin / kbox; p.conv_tiles_x = (W + 15) / 16; p.conv_tiles_y = (H + 7) / 8;
    p.num_m_tiles = T * p.conv_tiles_x * p.conv_tiles_y;
    p.num_n_tiles = 0;
    p.cluster = 1;
    const bool wide = (Cout % 256 == 0);
    {
        static const char *env = getenv("VNEXT_GEMM_CLUSTER");
        const int want = env ? atoi(env) : 1;
        if (want == 2 && p.num_m_tiles * ((Cout + (wide ? 255 : 127)) / (wide ? 256 : 128)) >= 2 * num_sms()) p.cluster = 2;
    }
    CUtensorMap ta, tb, to;
    int rc = make_tmap_nhwc(&ta, in_dt, x, T, Hin, Win, Cin, 8, 16, kbox, stride);
    if (rc) return rc;
    rc = make_tmap_2d(&tb, in_dt, w, Cout, 9 * Cin, 9 * Cin, (wide ? 256 : 128) / p.cluster, kbox);
    if (rc) return rc;
    rc = make_tmap_nhwc(&to, out_dt, out, T, H, W, Cout, 2, 16, 32);
    if (rc) return rc;
    return launch_any(wide, in_dt, out_dt, ta, tb, to, to, p, stream);
}

// 7x7 stride-2 padding-3 convolution of a 3-channel image (the ResNet stem, detectron2/modeling/backbone/resnet.py:355-359) as
// an implicit GEMM with K = 7 filter rows x 32.  `xpad`: [T, Hp + 6, Wp + 6, 4] fp32 — the normalised frame with a zero border
// of 3 pixels and a zero 4th channel (written by vnext_stem_conv7x7's preprocess pass), so a pixel is 16 B and a filter row's
// window of output pixel ox is the 32 floats starting at padded column 2*ox: tensor map dims {32, Wo, Hp + 6, T} with strides
// {32 B, row pitch, image pitch} — rows overlap, every address is inside the buffer, no out-of-bounds fill needed.
// `w`: [Cout, 7 * 32] (ky, kx padded to 8, channel padded to 4).
int stem_conv_device(int out_dt, const float *xpad, const float *w, const float *bias, void *out, int T, int Hp, int Wp, int Cout,
                     int relu, cudaStream_t stream) {
    const int Ho = Hp / 2, Wo = Wp / 2;
    GemmParams p;
    p.M = T * Ho * Wo; p.N = Cout; p.K = 7 * 32; p.bias = bias; p.bias2d = nullptr; p.period = 1; p.ldb2 = 0;
    p.b2_tma = 0; p.pdl = 0; p.f16_in = p.f16_out = 0; p.msda_ref = nullptr;
    p.row_mask = nullptr; p.mask_cols = 0; p.residual = nullptr; p.gamma = p.beta = nullptr; p.out = static_cast<float *>(out);
    p.ldo = Cout; p.ldr = 0; p.flags = relu ? EPI_RELU : 0; p.eps = 0.f; p.vec_ok = 1; p.tma_epi = 1;
    p.conv = 2; p.conv_cin_blocks = 1; p.conv_stride = 2; p.conv_tiles_x = (Wo + 127) / 128; p.conv_tiles_y = Ho;
    p.num_m_tiles = T * p.conv_tiles_x * p.conv_tiles_y;
    p.num_n_tiles = 0;
    p.cluster = 1;
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VNEXT_ECUDA;
    }
    CUtensorMap ta, tb, to;
    {
        const cuuint64_t rowpitch = (cuuint64_t)(Wp + 6) * 16, imgpitch = rowpitch * (Hp + 6);
        cuuint64_t dims[4] = {32, (cuuint64_t)Wo, (cuuint64_t)(Hp + 6), (cuuint64_t)T};
        cuuint64_t strides[3] = {32, rowpitch, imgpitch};
        cuuint32_t box[4] = {32u, 128u, 1u, 1u};
        cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
        CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(xpad), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled (stem im2col view) failed (%d)", (int)r);
            return VNEXT_ECUDA;
        }
    }
    int rc = make_tmap_2d(&tb, VNEXT_F32, w, Cout, 7 * 32, 7 * 32, 128, 32);
    if (rc) return rc;
    rc = make_tmap_nhwc(&to, out_dt, out, T, Ho, Wo, Cout, 1, 32, 32);
    if (rc) return rc;
    return launch_any(false, VNEXT_F32, out_dt, ta, tb, to, to, p, stream);
}

}  // namespace vnext

using namespace vnext;

static int conv3x3_checked(const char *who, int in_dt, int out_dt, const void *x, const void *w, const float *bias, void *out,
                           int T, int H, int W, int Cin, int Cout, int stride, int relu, void *stream) {
    VNEXT_CHECK_ARG(stride == 1 || stride == 2, "%s: stride must be 1 or 2 (got %d)", who, stride);
    VNEXT_CHECK_ARG((in_dt == VNEXT_F32 || in_dt == VNEXT_BF16) && (out_dt == VNEXT_F32 || out_dt == VNEXT_BF16),
                    "%s: dtypes must be fp32 (0) or bf16 (2)", who);
    VNEXT_CHECK_ARG(T >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "%s: bad shape", who);
    const int kb = in_dt == VNEXT_BF16 ? 64 : 32, oa = out_dt == VNEXT_BF16 ? 8 : 4;
    VNEXT_CHECK_ARG(Cin % kb == 0 && Cout % oa == 0, "%s: needs Cin %% %d == 0 and Cout %% %d == 0 (got %d, %d)", who, kb, oa, Cin, Cout);
    if (T == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(x && w && out && al16(x) && al16(w) && al16(out) && (!bias || al16(bias)), "%s: null or unaligned pointer", who);
    VNEXT_CHECK_ARG((long long)T * H * W < (1ll << 31), "%s: too many pixels", who);
    return conv3x3_tc_device(in_dt, out_dt, x, w, bias, out, T, H, W, Cin, Cout, stride, relu, static_cast<cudaStream_t>(stream));
}

extern "C" int vnext_conv3x3_tf32(const float *x, const float *weight_khwc, const float *bias, float *out, int T, int H,
                                  int W, int Cin, int Cout, int relu, void *stream) {
    return conv3x3_checked("vnext_conv3x3_tf32", VNEXT_F32, VNEXT_F32, x, weight_khwc, bias, out, T, H, W, Cin, Cout, 1, relu, stream);
}

extern "C" int vnext_conv3x3_tc(int in_dtype, int out_dtype, const void *x, const void *weight_khwc, const float *bias,
                                void *out, int T, int H, int W, int Cin, int Cout, int relu, void *stream) {
    return conv3x3_checked("vnext_conv3x3_tc", in_dtype, out_dtype, x, weight_khwc, bias, out, T, H, W, Cin, Cout, 1, relu, stream);
}

extern "C" int vnext_conv3x3_strided_tc(int in_dtype, int out_dtype, const void *x, const void *weight_khwc, const float *bias,
                                        void *out, int T, int H, int W, int Cin, int Cout, int stride, int relu, void *stream) {
    return conv3x3_checked("vnext_conv3x3_strided_tc", in_dtype, out_dtype, x, weight_khwc, bias, out, T, H, W, Cin, Cout, stride,
                           relu, stream);
}

static int linear_checked(const char *who, int in_dt, int out_dt, const void *a, int lda, const void *weight, const float *bias,
                          const float *bias2d, int bias_period, int ldb2, const unsigned char *row_mask, int mask_cols,
                          const void *residual, int ldr, const float *ln_gamma, const float *ln_beta, float ln_eps, void *out,
                          int ldo, int M, int N, int K, int flags, void *stream) {
    VNEXT_CHECK_ARG((in_dt == VNEXT_F32 || is16(in_dt)) && (out_dt == VNEXT_F32 || is16(out_dt)),
                    "%s: dtypes must be fp32 (0), bf16 (2) or fp16 (3)", who);
    VNEXT_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "%s: bad sizes M=%d N=%d K=%d", who, M, N, K);
    VNEXT_CHECK_ARG(flags >= 0 && flags < 0x800 && !(flags & 0x0f0), "%s: unknown epilogue flags %d", who, flags);
    if (M == 0 || N == 0) return VNEXT_OK;
    VNEXT_CHECK_ARG(a && weight && out, "%s: null pointer", who);
    const int ie = is16(in_dt) ? 8 : 4;
    VNEXT_CHECK_ARG(al16(a) && al16(weight) && lda % ie == 0 && K % ie == 0,
                    "%s: a / weight must be 16-byte aligned with lda and K multiples of %d (TMA)", who, ie);
    VNEXT_CHECK_ARG(!(flags & EPI_RESIDUAL) || residual, "%s: residual flag without residual pointer", who);
    VNEXT_CHECK_ARG(!bias2d || bias_period > 0, "%s: bias2d needs a positive period", who);
    VNEXT_CHECK_ARG(!row_mask || bias_period > 0, "%s: row_mask needs a positive period", who);
    LinearArgs x{in_dt, out_dt, a, lda, weight, bias, bias2d, bias_period, ldb2, row_mask, mask_cols, residual, ldr,
                 ln_gamma, ln_beta, ln_eps, out, ldo, M, N, K, flags};
    return linear_tc_device(x, static_cast<cudaStream_t>(stream));
}

extern "C" int vnext_linear_tf32(const float *a, int lda, const float *weight, const float *bias, const float *bias2d,
                                 int bias_period, int ldb2, const unsigned char *row_mask, int mask_cols,
                                 const float *residual, int ldr, const float *ln_gamma, const float *ln_beta,
                                 float ln_eps, float *out, int ldo, int M, int N, int K, int flags, void *stream) {
    return linear_checked("vnext_linear_tf32", VNEXT_F32, VNEXT_F32, a, lda, weight, bias, bias2d, bias_period, ldb2, row_mask,
                          mask_cols, residual, ldr, ln_gamma, ln_beta, ln_eps, out, ldo, M, N, K, flags, stream);
}

extern "C" int vnext_linear_tc(int in_dtype, int out_dtype, const void *a, int lda, const void *weight, const float *bias,
                               const float *bias2d, int bias_period, int ldb2, const unsigned char *row_mask, int mask_cols,
                               const void *residual, int ldr, const float *ln_gamma, const float *ln_beta, float ln_eps,
                               void *out, int ldo, int M, int N, int K, int flags, void *stream) {
    return linear_checked("vnext_linear_tc", in_dtype, out_dtype, a, lda, weight, bias, bias2d, bias_period, ldb2, row_mask,
                          mask_cols, residual, ldr, ln_gamma, ln_beta, ln_eps, out, ldo, M, N, K, flags, stream);
}

extern "C" int vnext_linear_msda_front(int in_dtype, const void *a, int lda, const void *weight, const float *bias2d, int bias_period,
                                       int ldb2, const float *reference_points, const float *level_wh, float *out, int ldo, int M,
                                       int K, void *stream) {
    VNEXT_CHECK_ARG(in_dtype == VNEXT_F32 || in_dtype == VNEXT_BF16, "vnext_linear_msda_front: dtype must be fp32 (0) or bf16 (2)");
    VNEXT_CHECK_ARG(M >= 0 && K > 0 && bias_period > 0, "vnext_linear_msda_front: bad sizes M=%d K=%d period=%d", M, K, bias_period);
    if (M == 0) return VNEXT_OK;
    const int ie = in_dtype == VNEXT_BF16 ? 8 : 4;
    VNEXT_CHECK_ARG(a && weight && bias2d && reference_points && level_wh && out && al16(a) && al16(weight) && al16(out) &&
                        al16(reference_points) && lda % ie == 0 && K % ie == 0 && ldo % 4 == 0,
                    "vnext_linear_msda_front: null or unaligned pointer (16-byte alignment, lda / K multiples of %d)", ie);
    LinearArgs x{in_dtype, VNEXT_F32, a, lda, weight, nullptr, bias2d, bias_period, ldb2, nullptr, 0, nullptr, 0,
                 nullptr, nullptr, 0.f, out, ldo, M, 384, K, 0};
    x.msda_ref = reference_points;
    x.msda_wh = level_wh;
    return linear_tc_device(x, static_cast<cudaStream_t>(stream));
}
