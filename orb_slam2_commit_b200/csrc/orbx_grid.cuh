// orbx_grid.cuh — Frame::mGrid / KeyFrame::mGrid (Frame.cc:254-271, 446-460; KeyFrame.cc:52-57 copies it) rebuilt in
// shared memory as a CSR by one CTA: cell (posX, posY) -> keypoint indices in ascending order, cells numbered
// posX * 48 + posY so that the rows cy0..cy1 of one cell column are ONE contiguous CSR range. Shared by the matchers that
// walk GetFeaturesInArea windows (orbx_project.cu, orbx_match.cu).
#pragma once
#include "orbx_internal.cuh"

#define GRID_COLS 64                                           // FRAME_GRID_COLS (Frame.h)
#define GRID_ROWS 48                                           // FRAME_GRID_ROWS
#define GRID_CELLS (GRID_COLS * GRID_ROWS)

struct GridKp { float x, y; int octave; };

struct GridSmem {                                              // carved from dynamic shared memory by grid_carve()
    GridKp* kp;                                                // [n]
    unsigned short* order;                                     // [n] keypoint indices sorted by cell
    unsigned short* cstart;                                    // [GRID_CELLS + 1]
    unsigned short* cfill;                                     // [GRID_CELLS] counts, then fill cursors
};

// bytes of dynamic shared memory grid_carve() consumes for n keypoints
static inline size_t grid_smem_bytes(size_t n)
{
    n = n ? n : 1;
    return n * sizeof(GridKp) + ((n + 1) & ~(size_t)1) * 2 + (size_t)(2 * GRID_CELLS + 4) * 2;
}

__device__ __forceinline__ unsigned char* grid_carve(unsigned char* base, int n, GridSmem& g)
{
    g.kp = reinterpret_cast<GridKp*>(base);
    g.order = reinterpret_cast<unsigned short*>(g.kp + n);
    g.cstart = g.order + ((n + 1) & ~1);
    g.cfill = g.cstart + GRID_CELLS + 2;
    return reinterpret_cast<unsigned char*>(g.cfill + GRID_CELLS + 2);
}

// Frame::PosInGrid (Frame.cc:446-460): cell id or -1
__device__ __forceinline__ int grid_cell_of(float x, float y, float minX, float minY, float invW, float invH)
{
    const int posX = (int)roundf(__fmul_rn(__fsub_rn(x, minX), invW));
    const int posY = (int)roundf(__fmul_rn(__fsub_rn(y, minY), invH));
    return (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) ? -1 : posX * GRID_ROWS + posY;
}

// Counting sort of the keypoints by cell (two 16-bit counters per word), then a per-cell insertion sort so that every
// cell lists ascending indices like AssignFeaturesToGrid. All threads of a 512-thread CTA must call it; s_w = 17 ints of
// static shared memory. Ends with a barrier.
__device__ __forceinline__ void grid_build(const OrbxKp28* __restrict__ kps, int n, float minX, float minY, float invW,
                                           float invH, const GridSmem& g, int* s_w)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int c = tid; c < GRID_CELLS; c += blockDim.x) g.cfill[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const OrbxKp28 k = kps[i];
        GridKp e; e.x = k.x; e.y = k.y; e.octave = k.octave;
        g.kp[i] = e;
        const int c = grid_cell_of(k.x, k.y, minX, minY, invW, invH);
        if (c >= 0) atomicAdd(reinterpret_cast<unsigned*>(g.cfill) + (c >> 1), (c & 1) ? 0x10000u : 1u);
    }
    __syncthreads();
    {   // exclusive scan of the cell counts: 6 cells per thread (512 x 6 = 3072), then a block scan of the per-thread sums
        const int c0 = tid * 6;
        int loc[6], sum = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) { loc[j] = sum; sum += g.cfill[c0 + j]; }
        int x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            const int t = lane < 16 ? s_w[lane] : 0;
            int z = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, z, o); if (lane >= o) z += y; }
            if (lane < 16) s_w[lane] = z - t;
            if (lane == 15) s_w[16] = z;
        }
        __syncthreads();
        const int base = s_w[wid] + x - sum;
#pragma unroll
        for (int j = 0; j < 6; j++) { g.cstart[c0 + j] = (unsigned short)(base + loc[j]); g.cfill[c0 + j] = (unsigned short)(base + loc[j]); }
        if (tid == 0) g.cstart[GRID_CELLS] = (unsigned short)s_w[16];
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const GridKp k = g.kp[i];
        const int c = grid_cell_of(k.x, k.y, minX, minY, invW, invH);
        if (c >= 0) {
            const unsigned old = atomicAdd(reinterpret_cast<unsigned*>(g.cfill) + (c >> 1), (c & 1) ? 0x10000u : 1u);
            g.order[(c & 1) ? (old >> 16) : (old & 0xffffu)] = (unsigned short)i;
        }
    }
    __syncthreads();
    // the atomics filled every cell in arbitrary order: the reference's cells hold ascending indices
    for (int c = tid; c < GRID_CELLS; c += blockDim.x) {
        const int b0 = g.cstart[c], b1 = g.cstart[c + 1];
        for (int i = b0 + 1; i < b1; i++) {
            const unsigned short v = g.order[i];
            int j = i - 1;
            while (j >= b0 && g.order[j] > v) { g.order[j + 1] = g.order[j]; j--; }
            g.order[j + 1] = v;
        }
    }
    __syncthreads();
}

// Cell rectangle of GetFeaturesInArea (Frame.cc:394-408 / KeyFrame.cc:713-727); false = the window misses the grid
__device__ __forceinline__ bool grid_window(float x, float y, float r, float minX, float minY, float invW, float invH,
                                            int& cx0, int& cx1, int& cy0, int& cy1)
{
    cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
    cx1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
    cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
    cy1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
    return !(cx0 >= GRID_COLS || cx1 < 0 || cy0 >= GRID_ROWS || cy1 < 0);
}

__device__ __forceinline__ int grid_hamming(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}
