// orbx_describe.cu — orientation + 7x7 Gaussian + steered rBRIEF, one warp per keypoint (replaces
// computeOrientation/IC_Angle ORBextractor.cc:77-105,492-499, the per-level cv::GaussianBlur :1188-1190 and
// computeOrbDescriptor :110-152, plus the coordinate scaling / output packing of operator() :1194-1209).
//
// One elected lane has the TMA unit copy two boxes around the keypoint into the warp's shared-memory slots
// (cp.async.bulk.tensor through the level's tensor maps, completion on the warp's mbarrier): 48x31 bytes of the RAW level for
// the orientation and 64x37 bytes of the BLURRED level (orbx_blur.cu) for the tests; then
//   * IC_Angle: lane = column u in [-15,15], integer moments, warp-shuffle reduce, cv::fastAtan2 polynomial
//     evaluated with un-contracted f32 mul/add (bit-equal to OpenCV's scalar path);
//   * rBRIEF: lane i builds descriptor byte i (8 tests); sample = center + cvRound(x*b+y*a, x*a-y*b) with
//     un-contracted f32 and round-half-even, one byte load from the blurred box; cos/sin are the glibc 2.39 cosf/sinf
//     polynomials in f64 (bit-equal to the host libm the oracle was pinned against, tests/golden/sincos.json).
// Round 1 blurred the 43x43 patch inside this kernel (no blurred pyramid in HBM): 750 of its 1500 warp-instructions per
// keypoint; the dense blur costs a third of that per frame.
#include "orbx_internal.cuh"
#include "orbx_tma.cuh"
#include <cstdlib>

#ifndef DESC_WARPS
#define DESC_WARPS 4         // warps (= keypoints) per CTA
#endif
#define RAW_W 48         // raw box: columns kx-15 .. kx+15 (+ up to 15 of alignment: the box origin is 16-byte aligned), rows ky-15 .. ky+15
#define RAW_H 31
#define BLR_W 64         // blurred box: columns kx-18 .. kx+18 (+ alignment), rows ky-18 .. ky+18
#define BLR_H 37
#define RAW_SLOT 1536    // 48 x 31 = 1488, padded to a multiple of 128 bytes (TMA destination alignment)
#define BLR_SLOT 2432    // 64 x 37 = 2368

__device__ uint32_t g_pattern32[256];    // the 512 (x,y) int8 sample points, four bytes per word
__constant__ int c_umax[16];

static const signed char h_pattern[1024] = {
#include "orb_pattern.inc"
};

void orbx_upload_constants()
{
    // IC_Angle disc half-widths (ORBextractor.cc:473-489)
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    cudaMemcpyToSymbol(g_pattern32, h_pattern, sizeof(h_pattern));
    cudaMemcpyToSymbol(c_umax, umax, sizeof(umax));
}

__device__ __forceinline__ float dev_fast_atan2(const float y, const float x)
{
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float eps = 2.2204460492503131e-16f;   // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

// glibc 2.39 x86_64 sincosf (sysdeps/ieee754/flt-32/s_sincosf.h, FMA build), medium-range path |x| < 120.
__device__ __forceinline__ double sc_cos_poly(const double x2, const double sg)
{
    const double c0 = 1.0, c1 = __longlong_as_double(0xbfdffffffd0c621cLL), c2 = __longlong_as_double(0x3fa55553e1068f19LL);
    const double c3 = __longlong_as_double(0xbf56c087e89a359dLL), c4 = __longlong_as_double(0x3ef99343027bf8c3LL);
    const double x4 = __dmul_rn(x2, x2);
    const double q2 = __fma_rn(sg * c4, x2, sg * c3);
    const double q1 = __fma_rn(sg * c1, x2, sg * c0);
    const double x6 = __dmul_rn(x2, x4);
    const double q = __fma_rn(x4, sg * c2, q1);
    return __fma_rn(q2, x6, q);
}
__device__ __forceinline__ double sc_sin_poly(const double x, const double x2)
{
    const double s1 = __longlong_as_double(0xbfc555545995a603LL), s2 = __longlong_as_double(0x3f81107605230bc4LL);
    const double s3 = __longlong_as_double(0xbf2994eb3774cf24LL);
    const double t1 = __fma_rn(s3, x2, s2);
    const double x3 = __dmul_rn(x2, x);
    const double x7 = __dmul_rn(x2, x3);
    const double s = __fma_rn(x3, s1, x);
    return __fma_rn(t1, x7, s);
}
__device__ __forceinline__ void dev_sincosf(const float y, float* sn, float* cs)
{
    const unsigned top = (__float_as_uint(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top <= 0x3f3) {
        const double x2 = __dmul_rn(x, x);
        if (top <= 0x397) { *cs = 1.0f; *sn = y; return; }
        *cs = __double2float_rn(sc_cos_poly(x2, 1.0));
        *sn = __double2float_rn(sc_sin_poly(x, x2));
        return;
    }
    const double hpi_inv = __longlong_as_double(0x41645f306dc9c883LL), hpi = __longlong_as_double(0x3ff921fb54442d18LL);
    const double r = __dmul_rn(x, hpi_inv);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __fma_rn(-(double)n, hpi, x);
    const double x2 = __dmul_rn(x, x);
    const double sg = (n & 2) ? -1.0 : 1.0;                      // table 1 = negated cosine coefficients
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   // sign[n & 3] = {1,-1,-1,1}
    const double pc = sc_cos_poly(x2, sg);
    const double ps = sc_sin_poly(__dmul_rn(x, sign), x2);
    // cosf: (n&1)==0 -> cosine polynomial, else sine polynomial; sinf: the other way round
    *cs = __double2float_rn((n & 1) ? ps : pc);
    *sn = __double2float_rn((n & 1) ? pc : ps);
}

// A warp takes `kpw` consecutive keypoint ordinals of a frame: the pattern (converted to f32 once), the level totals and the
// barrier set-up are paid once per warp, and the boxes of keypoint j+1 are in flight (second pair of slots) while keypoint j is
// processed.
__global__ void __launch_bounds__(DESC_WARPS * 32) describe_kernel(OrbxFrameLayout L, const __grid_constant__ OrbxTmaps maps_raw,
                                                                    const __grid_constant__ OrbxTmaps maps_blur,
                                                                    OrbxKp28* __restrict__ kps, uint8_t* __restrict__ desc, int cap,
                                                                    int* __restrict__ nkp, int kpw)
{
    __shared__ __align__(128) uint8_t s_raw[DESC_WARPS][2][RAW_SLOT];
    __shared__ __align__(128) uint8_t s_blr[DESC_WARPS][2][BLR_SLOT];
    __shared__ __align__(8) unsigned long long s_bar[DESC_WARPS][2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // this lane's 16 sample points (32 int8 = two 128-bit words); no block-level barrier anywhere in this kernel
    const uint4 pat_lo = __ldg(reinterpret_cast<const uint4*>(g_pattern32) + 2 * lane);
    const uint4 pat_hi = __ldg(reinterpret_cast<const uint4*>(g_pattern32) + 2 * lane + 1);
    const int frame = blockIdx.y;
    const int ord0 = (blockIdx.x * DESC_WARPS + wid) * kpw;      // first keypoint ordinal of this warp (level-major inside the frame)
    const int* cnt = L.lvl_kp_count + (size_t)frame * L.nlevels;
    // lane l holds the count of level l, a warp scan gives the running totals (nlevels <= 16)
    const int c_lvl = lane < L.nlevels ? cnt[lane] : 0;
    int incl = c_lvl;
#pragma unroll
    for (int o = 1; o < ORBX_MAX_LEVELS; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, ORBX_MAX_LEVELS - 1);   // the scan spans 16 lanes
    if (ord0 == 0 && lane == 0) nkp[frame] = total;
    const int nk = min(kpw, min(total, cap) - ord0);
    if (nk <= 0) return;                                 // warp-uniform
    float patx[16], paty[16];
    {
        const uint32_t patw[8] = {pat_lo.x, pat_lo.y, pat_lo.z, pat_lo.w, pat_hi.x, pat_hi.y, pat_hi.z, pat_hi.w};
#pragma unroll
        for (int i = 0; i < 16; i++) {                   // (x0, y0, x1, y1) of test i / 2 as four int8
            patx[i] = (float)(signed char)(patw[i >> 1] >> (16 * (i & 1)));
            paty[i] = (float)(signed char)(patw[i >> 1] >> (16 * (i & 1) + 8));
        }
    }
    // IC_Angle: this lane's row of the disc as byte masks of its eight words (row v = lane - 15 spans |u| <= umax[|v|]; lane 31 idle)
    unsigned disc[8];
    {
        const int d = lane < 31 ? c_umax[lane < 15 ? 15 - lane : lane - 15] : -1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned mk = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int u = 4 * k - 15 + j;
                if ((u < 0 ? -u : u) <= d) mk |= 0xffu << (8 * j);
            }
            disc[k] = mk;
        }
    }
    const uint32_t bar0 = orbx_smem_addr(&s_bar[wid][0]);
    if (lane == 0) { orbx_mbar_init(bar0, 1); orbx_mbar_init(bar0 + 8, 1); }
    __syncwarp();                                        // the barriers are initialised before anyone arms or waits on them

    // level and packed (x, y, score) of ordinal `ord`; one elected lane starts the copies of its two boxes into slot `s`:
    // each box starts at the 16-byte aligned column at or left of its window
    auto stage = [&](const int ord, const int s, int& level, uint32_t& pk) {
        const unsigned m = __ballot_sync(0xffffffffu, ord < incl);
        level = __ffs(m) - 1;
        const int k = ord - __shfl_sync(0xffffffffu, incl - c_lvl, level);
        pk = L.lvl_kp[(size_t)frame * L.kp_cap_total + L.lvl_kp_off[level] + k];
        if (lane == 0) {
            const int kx = pk & 0xfff, ky = (pk >> 12) & 0xfff;
            const uint32_t bar = bar0 + 8 * s;
            orbx_fence_proxy_async();                    // the slot's previous reader (generic proxy) is done: __syncwarp at the loop end
            orbx_mbar_expect_tx(bar, RAW_W * RAW_H + BLR_W * BLR_H);
            orbx_tma_load_3d(orbx_smem_addr(s_raw[wid][s]), &maps_raw.m[level], (kx - 15 + ORBX_XOFF) & ~15, ky - 15 + ORBX_EDGE, L.frame0 + frame, bar);
            orbx_tma_load_3d(orbx_smem_addr(s_blr[wid][s]), &maps_blur.m[level], (kx - 18 + ORBX_XOFF) & ~15, ky - 18 + ORBX_EDGE, L.frame0 + frame, bar);
        }
    };
    int level, level_n = 0;
    uint32_t pk, pk_n = 0;
    stage(ord0, 0, level, pk);
    for (int j = 0; j < nk; j++) {
    const int s = j & 1;
    if (j + 1 < nk) stage(ord0 + j + 1, s ^ 1, level_n, pk_n);
    orbx_mbar_wait(bar0 + 8 * s, (j >> 1) & 1);
    const int ord = ord0 + j;
    const int kx = pk & 0xfff, ky = (pk >> 12) & 0xfff, score = pk >> 24;
    const int xr = kx - 15 + ORBX_XOFF, xb = kx - 18 + ORBX_XOFF;

    // ---- IC_Angle on the un-blurred level: lane = row v of the disc (31 rows), the row's 31 bytes as eight words brought
    // to byte 0 by funnel shifts (the box row starts xr & 3 bytes before a word boundary, warp-uniform), the bytes outside
    // the disc masked off; m10 += sum_u u * I by one mixed-sign dp4a per word against the packed column offsets, the row
    // sum by one dp4a against ones, m01 += v * row sum. Integer sums: any order gives the reference's result
    // (ORBextractor.cc:91-102 pairs rows +v and -v). 65 instead of 134 instructions per keypoint.
    int m10 = 0, m01;
    {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(s_raw[wid][s]) + ((xr & 15) >> 2) + 12 * min(lane, 30);
        const unsigned sh8 = 8 * (xr & 3);
        uint32_t w[9];
#pragma unroll
        for (int k = 0; k < 9; k++) w[k] = rw[k];
        unsigned rs = 0, rs2 = 0;
        int m10b = 0;                                      // two accumulator chains each: the dot products are dependent otherwise
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned x = __funnelshift_r(w[k], w[k + 1], sh8) & disc[k];
            const int u0 = 4 * k - 15;
            const unsigned wt = (unsigned)(u0 & 0xff) | ((unsigned)((u0 + 1) & 0xff) << 8) | ((unsigned)((u0 + 2) & 0xff) << 16) | ((unsigned)((u0 + 3) & 0xff) << 24);
            if (k & 1) { asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10b) : "r"(x), "r"(wt)); rs2 = __dp4a(x, 0x01010101u, rs2); }
            else { asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(x), "r"(wt)); rs = __dp4a(x, 0x01010101u, rs); }
        }
        m10 += m10b;
        m01 = (lane - 15) * (int)(rs + rs2);
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);           // redux.sync: one instruction per warp sum
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = dev_fast_atan2((float)m01, (float)m10);

    // ---- steered rBRIEF: lane i -> descriptor byte i
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float a, b;
    dev_sincosf(__fmul_rn(angle, factorPI), &b, &a);
    // cvRound by the 1.5 * 2^23 trick (round-half-even like F2I, but on the FMA pipe): bits(v + M) = bits(M) + rint(v) for
    // |v| < 2^22, so iy * BLR_W + ix = bits_y * BLR_W + bits_x - (BLR_W + 1) * bits(M); the constant and the offset of the
    // blurred pixel at the keypoint inside the slot are folded into one base (32-bit wrap-around arithmetic)
    const float MAGIC = 12582912.f;
    const uint8_t* blr = s_blr[wid][s];
    const unsigned base = (unsigned)((xb & 15) + 18 * BLR_W + 18) - (unsigned)(BLR_W + 1) * 0x4B400000u;
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        int smp[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = patx[2 * t + e], py = paty[2 * t + e];
            const unsigned by = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), MAGIC));
            const unsigned bx = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), MAGIC));
            smp[e] = blr[by * BLR_W + bx + base];
        }
        val |= (smp[0] < smp[1]) << t;
    }
    const size_t o = (size_t)frame * cap + ord;
    desc[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        const OrbxLevelGeom* __restrict__ g = L.lvl + level;
        OrbxKp28 kp;
        kp.x = level ? __fmul_rn((float)kx, g->scale) : (float)kx;
        kp.y = level ? __fmul_rn((float)ky, g->scale) : (float)ky;
        kp.size = g->kp_size;
        kp.angle = angle;
        kp.response = (float)score;
        kp.octave = level;
        kp.class_id = -1;
        kps[o] = kp;
    }
    __syncwarp();                                        // every lane is done with slot s before it is refilled
    level = level_n; pk = pk_n;
    }
}

void orbx_launch_describe(const OrbxFrameLayout& L, const OrbxTmaps& maps_raw, const OrbxTmaps& maps_blur, int nframes, OrbxKp28* d_kps,
                          uint8_t* d_desc, int cap, int* d_nkp, cudaStream_t st)
{
    static const int kpw_env = getenv("ORBX_DESC_KPW") ? atoi(getenv("ORBX_DESC_KPW")) : 0;
    // few frames: one keypoint per warp keeps every SM busy; batches: eight per warp amortise the per-warp set-up
    // (describe stage of 512 VGA frames: 0.583 / 0.508 / 0.486 / 0.477 ms with 1 / 2 / 4 / 8)
    const int kpw = kpw_env > 0 ? kpw_env : (nframes >= 8 ? 8 : 1);
    const int per_cta = DESC_WARPS * kpw;
    dim3 grid((L.kp_cap_total + per_cta - 1) / per_cta, nframes);
    describe_kernel<<<grid, DESC_WARPS * 32, 0, st>>>(L, maps_raw, maps_blur, d_kps, d_desc, cap, d_nkp, kpw);
}
