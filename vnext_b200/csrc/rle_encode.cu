// COCO run-length encoding of boolean masks on the device (the wire format the reference's evaluator produces
// right after the hot path — projects/IDOL/idol/data/ytvis_eval.py:174-214, projects/SeqFormer/seqformer/data/ytvis_eval.py:
// `mask_util.encode(np.array(mask[:, :, None], order="F"))`, i.e. pycocotools maskApi.c rleEncode + rleToString).
//
// A mask [H, W] (row-major bytes, 0 / non-zero) is read as the COLUMN-major pixel stream i = x*H + y.  Six small launches,
// all of fixed grid size (CUDA-graph friendly), all integer work bound by the mask reads (a thread walks 4 adjacent columns
// with one 4-byte load per row when W % 4 == 0; rows are coalesced across the warp):
//   1 count     unit = (column x, one of RLE_SEG row segments): #positions where the stream changes value
//   2 scan      per mask, exclusive scan of the unit counts (unit order x*RLE_SEG + seg == stream order) -> n_runs
//   3 emit      every unit writes the stream positions of its changes: positions[mask][0 .. n_runs)
//   4 lengths   count i = pos(i) - pos(i-1)  (pos(-1) = 0, pos(n_runs) = H*W);  chars(count_i - count_{i-2} [i > 2]) summed
//               per chunk of RLE_CHUNK counts
//   5 scan      one block: exclusive scan of all (mask, chunk) sums -> where every chunk's characters start in the single
//               output string buffer, char_offsets[mask], char_offsets[n_masks] = total
//   6 write     the characters (5-bit groups, LSB first, 0x20 = more, + 48)
// Masks' strings are concatenated: the host copies char_offsets (n_masks + 1 int64) and total bytes — KBs for real masks
// instead of H*W bytes per mask.
#include "common.cuh"

using namespace vnext;

namespace {

constexpr int RLE_SEG = 16;           // row segments per column
constexpr int RLE_COLS = 64;          // threads per CTA of the count / emit kernels (each owns 1 or 4 adjacent columns)
constexpr int RLE_CHUNK = 2048;       // counts per CTA of the string kernels (256 threads x 8)
constexpr int RLE_PER_THREAD = RLE_CHUNK / 256;

struct RleLayout {
    size_t positions, unit_offsets, n_runs, chunk_chars, chunk_offsets, total;
    int n_units, n_chunks;
};

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

RleLayout rle_layout(int n_masks, int H, int W) {
    RleLayout l;
    const size_t hw = (size_t)H * W;
    l.n_units = W * RLE_SEG;
    l.n_chunks = (int)((hw + 1 + RLE_CHUNK - 1) / RLE_CHUNK);
    size_t o = 0;
    l.positions = o;      o = align256(o + sizeof(uint32_t) * hw * n_masks);
    l.unit_offsets = o;   o = align256(o + sizeof(int) * (size_t)l.n_units * n_masks);
    l.n_runs = o;         o = align256(o + sizeof(int) * (size_t)n_masks);
    l.chunk_chars = o;    o = align256(o + sizeof(int) * (size_t)l.n_chunks * n_masks);
    l.chunk_offsets = o;  o = align256(o + sizeof(long long) * (size_t)l.n_chunks * n_masks);
    l.total = o;
    return l;
}

// Exclusive scan of one value per thread across the CTA (blockDim.x a multiple of 32, <= 1024); `total` = the CTA's sum.
// scratch: 33 elements of shared memory; safe to call again right after it returns.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *scratch, T &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const T s = lane < n_warps ? scratch[lane] : T(0);
        T si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += up;
        }
        scratch[lane] = si - s;
        if (lane == 31) scratch[32] = si;
    }
    __syncthreads();
    total = scratch[32];
    const T res = scratch[warp] + inc - v;
    __syncthreads();
    return res;
}

// A thread owns VEC adjacent columns (one 4-byte load per row when VEC == 4) of one row segment: VEC units.
template <int VEC>
struct RleUnits {
    int x, y0, y1;          // first column, row range
    bool prev[VEC];         // stream value just before each column's first pixel of the segment (0 before the very first pixel)
};

template <int VEC>
__device__ __forceinline__ void rle_load_row(const unsigned char *row, int x, int W, bool (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const uchar4 q = *reinterpret_cast<const uchar4 *>(row + x);          // W % 4 == 0: aligned, all four columns exist
        v[0] = q.x != 0; v[1] = q.y != 0; v[2] = q.z != 0; v[3] = q.w != 0;
    } else {
        v[0] = row[x] != 0;
    }
}

template <int VEC>
__device__ __forceinline__ bool rle_units(const unsigned char *mask, int H, int W, RleUnits<VEC> &u) {
    u.x = (blockIdx.x * RLE_COLS + threadIdx.x) * VEC;
    if (u.x >= W) return false;
    const int seg_h = (H + RLE_SEG - 1) / RLE_SEG;
    u.y0 = blockIdx.y * seg_h;
    u.y1 = min(H, u.y0 + seg_h);
#pragma unroll
    for (int j = 0; j < VEC; ++j) u.prev[j] = false;
    if (u.y0 < u.y1) {
        if (u.y0 > 0) {
            rle_load_row<VEC>(mask + (size_t)(u.y0 - 1) * W, u.x, W, u.prev);
        } else {            // top of a column: the stream comes from the bottom of the column to the left
            bool last[VEC];
            rle_load_row<VEC>(mask + (size_t)(H - 1) * W, u.x, W, last);
#pragma unroll
            for (int j = 1; j < VEC; ++j) u.prev[j] = last[j - 1];
            u.prev[0] = u.x > 0 ? mask[(size_t)(H - 1) * W + u.x - 1] != 0 : false;
        }
    }
    return true;
}

template <int VEC>
__global__ void __launch_bounds__(RLE_COLS) rle_count_kernel(const unsigned char *__restrict__ masks, int H, int W,
                                                              int *__restrict__ unit_offsets) {
    const unsigned char *mask = masks + (size_t)blockIdx.z * H * W;
    RleUnits<VEC> u;
    if (!rle_units<VEC>(mask, H, W, u)) return;
    int cnt[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) cnt[j] = 0;
#pragma unroll 4
    for (int y = u.y0; y < u.y1; ++y) {
        bool v[VEC];
        rle_load_row<VEC>(mask + (size_t)y * W, u.x, W, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            cnt[j] += v[j] != u.prev[j];
            u.prev[j] = v[j];
        }
    }
    int *out = unit_offsets + (size_t)blockIdx.z * W * RLE_SEG + (size_t)u.x * RLE_SEG + blockIdx.y;
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[j * RLE_SEG] = cnt[j];
}

__global__ void __launch_bounds__(1024) rle_scan_units_kernel(int *__restrict__ unit_offsets, int n_units, int *__restrict__ n_runs) {
    __shared__ int scratch[33];
    int *mine = unit_offsets + (size_t)blockIdx.x * n_units;
    const int per = (n_units + blockDim.x - 1) / blockDim.x;
    const int i0 = min(n_units, (int)threadIdx.x * per), i1 = min(n_units, i0 + per);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += mine[i];
    int total;
    int run = block_exclusive_scan(sum, scratch, total);
    for (int i = i0; i < i1; ++i) {
        const int c = mine[i];
        mine[i] = run;
        run += c;
    }
    if (threadIdx.x == 0) n_runs[blockIdx.x] = total;
}

template <int VEC>
__global__ void __launch_bounds__(RLE_COLS) rle_emit_kernel(const unsigned char *__restrict__ masks, int H, int W,
                                                             const int *__restrict__ unit_offsets, uint32_t *__restrict__ positions) {
    const unsigned char *mask = masks + (size_t)blockIdx.z * H * W;
    RleUnits<VEC> u;
    if (!rle_units<VEC>(mask, H, W, u)) return;
    const int *offs = unit_offsets + (size_t)blockIdx.z * W * RLE_SEG + (size_t)u.x * RLE_SEG + blockIdx.y;
    uint32_t *base = positions + (size_t)blockIdx.z * H * W;
    uint32_t *out[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) out[j] = base + offs[j * RLE_SEG];
#pragma unroll 4
    for (int y = u.y0; y < u.y1; ++y) {
        bool v[VEC];
        rle_load_row<VEC>(mask + (size_t)y * W, u.x, W, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (v[j] != u.prev[j]) *out[j]++ = (uint32_t)(u.x + j) * (uint32_t)H + (uint32_t)y;
            u.prev[j] = v[j];
        }
    }
}

// maskApi.c rleToString, one count: number of characters of x.
__device__ __forceinline__ int rle_char_count(int x) {
    int n = 0;
    bool more = true;
    while (more) {
        const int c = x & 0x1f;
        x >>= 5;
        more = (c & 0x10) ? x != -1 : x != 0;
        ++n;
    }
    return n;
}

// The values written for counts [i0, i0 + RLE_PER_THREAD) of one mask: delta_i = count_i - (i > 2 ? count_{i-2} : 0).
__device__ __forceinline__ void rle_deltas(const uint32_t *__restrict__ pos, int k, int hw, int i0, int (&delta)[RLE_PER_THREAD]) {
    int p[RLE_PER_THREAD + 3];           // pos(i0 - 3 .. i0 + RLE_PER_THREAD - 1)
#pragma unroll
    for (int j = 0; j < RLE_PER_THREAD + 3; ++j) {
        const int i = i0 - 3 + j;
        p[j] = i < 0 ? 0 : (i >= k ? hw : (int)pos[i]);
    }
#pragma unroll
    for (int j = 0; j < RLE_PER_THREAD; ++j) {
        const int i = i0 + j;
        const int cnt = p[j + 3] - p[j + 2];
        delta[j] = i > 2 ? cnt - (p[j + 1] - p[j]) : cnt;
    }
}

__global__ void __launch_bounds__(256) rle_chunk_chars_kernel(const uint32_t *__restrict__ positions, const int *__restrict__ n_runs,
                                                               int hw, int n_chunks, int *__restrict__ chunk_chars) {
    __shared__ int scratch[33];
    const int m = blockIdx.y, k = n_runs[m], n_counts = k + 1;
    const int base = blockIdx.x * RLE_CHUNK;
    if (base >= n_counts) {
        if (threadIdx.x == 0) chunk_chars[(size_t)m * n_chunks + blockIdx.x] = 0;
        return;
    }
    const int i0 = base + threadIdx.x * RLE_PER_THREAD;
    int len = 0;
    if (i0 < n_counts) {
        int delta[RLE_PER_THREAD];
        rle_deltas(positions + (size_t)m * hw, k, hw, i0, delta);
#pragma unroll
        for (int j = 0; j < RLE_PER_THREAD; ++j)
            if (i0 + j < n_counts) len += rle_char_count(delta[j]);
    }
    int total;
    block_exclusive_scan(len, scratch, total);
    if (threadIdx.x == 0) chunk_chars[(size_t)m * n_chunks + blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) rle_scan_chunks_kernel(const int *__restrict__ chunk_chars, int n_chunks, int n_masks,
                                                                long long *__restrict__ chunk_offsets,
                                                                long long *__restrict__ char_offsets) {
    __shared__ long long scratch[33];
    const int n = n_chunks * n_masks;
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int i0 = min(n, (int)threadIdx.x * per), i1 = min(n, i0 + per);
    long long sum = 0;
    for (int i = i0; i < i1; ++i) sum += chunk_chars[i];
    long long total;
    long long run = block_exclusive_scan(sum, scratch, total);
    for (int i = i0; i < i1; ++i) {
        chunk_offsets[i] = run;
        if (i % n_chunks == 0) char_offsets[i / n_chunks] = run;
        run += chunk_chars[i];
    }
    if (threadIdx.x == 0) char_offsets[n_masks] = total;
}

__global__ void __launch_bounds__(256) rle_write_chars_kernel(const uint32_t *__restrict__ positions, const int *__restrict__ n_runs,
                                                               int hw, int n_chunks, const long long *__restrict__ chunk_offsets,
                                                               unsigned char *__restrict__ chars, long long capacity) {
    __shared__ int scratch[33];
    const int m = blockIdx.y, k = n_runs[m], n_counts = k + 1;
    const int base = blockIdx.x * RLE_CHUNK;
    if (base >= n_counts) return;
    const int i0 = base + threadIdx.x * RLE_PER_THREAD;
    int delta[RLE_PER_THREAD];
    int len = 0;
    if (i0 < n_counts) {
        rle_deltas(positions + (size_t)m * hw, k, hw, i0, delta);
#pragma unroll
        for (int j = 0; j < RLE_PER_THREAD; ++j)
            if (i0 + j < n_counts) len += rle_char_count(delta[j]);
    }
    int total;
    const int before = block_exclusive_scan(len, scratch, total);
    if (i0 >= n_counts) return;
    long long o = chunk_offsets[(size_t)m * n_chunks + blockIdx.x] + before;
#pragma unroll
    for (int j = 0; j < RLE_PER_THREAD; ++j) {
        if (i0 + j >= n_counts) break;
        int x = delta[j];
        bool more = true;
        while (more) {
            int c = x & 0x1f;
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            if (more) c |= 0x20;
            if (o < capacity) chars[o] = (unsigned char)(c + 48);
            ++o;
        }
    }
}

}  // namespace

extern "C" long long vnext_rle_workspace_bytes(int n_masks, int H, int W) {
    if (n_masks <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)rle_layout(n_masks, H, W).total;
}

extern "C" int vnext_rle_encode(const unsigned char *masks, int n_masks, int H, int W, void *workspace, long long workspace_bytes,
                                long long *char_offsets, unsigned char *chars, long long chars_capacity, void *stream) {
    VNEXT_CHECK_ARG(n_masks >= 0 && n_masks <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 31) - 1,
                    "vnext_rle_encode: needs 0 <= n_masks <= 65535 and 0 < H*W < 2^31 - 1 (n_masks %d, H %d, W %d)", n_masks, H, W);
    VNEXT_CHECK_ARG(char_offsets && chars_capacity >= 0 && (chars || chars_capacity == 0), "vnext_rle_encode: null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_masks == 0) {
        VNEXT_CUDA(cudaMemsetAsync(char_offsets, 0, sizeof(long long), st));
        return VNEXT_OK;
    }
    const RleLayout l = rle_layout(n_masks, H, W);
    VNEXT_CHECK_ARG(masks && workspace && workspace_bytes >= (long long)l.total && ((uintptr_t)workspace & 255) == 0,
                    "vnext_rle_encode: null / unaligned (256 B) pointer, or workspace of %lld bytes < vnext_rle_workspace_bytes() = %lld",
                    workspace_bytes, (long long)l.total);
    char *ws = static_cast<char *>(workspace);
    uint32_t *positions = reinterpret_cast<uint32_t *>(ws + l.positions);
    int *unit_offsets = reinterpret_cast<int *>(ws + l.unit_offsets);
    int *n_runs = reinterpret_cast<int *>(ws + l.n_runs);
    int *chunk_chars = reinterpret_cast<int *>(ws + l.chunk_chars);
    long long *chunk_offsets = reinterpret_cast<long long *>(ws + l.chunk_offsets);
    const int hw = H * W;
    const bool vec4 = W % 4 == 0 && ((uintptr_t)masks & 3) == 0;         // every row (and every mask) starts 4-byte aligned
    const int cols_per_cta = RLE_COLS * (vec4 ? 4 : 1);
    const dim3 unit_grid((W + cols_per_cta - 1) / cols_per_cta, RLE_SEG, n_masks), chunk_grid(l.n_chunks, n_masks);
    if (vec4) rle_count_kernel<4><<<unit_grid, RLE_COLS, 0, st>>>(masks, H, W, unit_offsets);
    else rle_count_kernel<1><<<unit_grid, RLE_COLS, 0, st>>>(masks, H, W, unit_offsets);
    rle_scan_units_kernel<<<n_masks, 1024, 0, st>>>(unit_offsets, l.n_units, n_runs);
    if (vec4) rle_emit_kernel<4><<<unit_grid, RLE_COLS, 0, st>>>(masks, H, W, unit_offsets, positions);
    else rle_emit_kernel<1><<<unit_grid, RLE_COLS, 0, st>>>(masks, H, W, unit_offsets, positions);
    rle_chunk_chars_kernel<<<chunk_grid, 256, 0, st>>>(positions, n_runs, hw, l.n_chunks, chunk_chars);
    rle_scan_chunks_kernel<<<1, 1024, 0, st>>>(chunk_chars, l.n_chunks, n_masks, chunk_offsets, char_offsets);
    rle_write_chars_kernel<<<chunk_grid, 256, 0, st>>>(positions, n_runs, hw, l.n_chunks, chunk_offsets, chars, chars_capacity);
    count_launch(6);
    VNEXT_CUDA(cudaGetLastError());
    return VNEXT_OK;
}
