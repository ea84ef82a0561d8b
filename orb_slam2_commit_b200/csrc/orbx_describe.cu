// orbx_describe.cu — orientation + 7x7 Gaussian + steered rBRIEF, one warp per keypoint (replaces
// computeOrientation/IC_Angle ORBextractor.cc:77-105,492-499, the per-level cv::GaussianBlur :1188-1190 and
// computeOrbDescriptor :110-152, plus the coordinate scaling / output packing of operator() :1194-1209).
//
// The warp stages the 43x43 raw patch around the keypoint in shared memory (aligned 32-bit loads), then
//   * IC_Angle: lane = column u in [-15,15], integer moments, warp-shuffle reduce, cv::fastAtan2 polynomial
//     evaluated with un-contracted f32 mul/add (bit-equal to OpenCV's scalar path);
//   * blur on demand: the reference blurs the whole level and then reads 512 points per keypoint; here the
//     horizontal pass of OpenCV's fixed-point kernel [18,34,48,56,48,34,18]/256 is applied to the patch (exact in
//     16 bits) and the vertical pass ((sum + 2^15) >> 16) only at the 16 sample points each lane needs. Border
//     handling is the level's own REFLECT_101, which is exactly what the 19-px apron in HBM holds;
//   * rBRIEF: lane i builds descriptor byte i (8 tests); sample = center + cvRound(x*b+y*a, x*a-y*b) with
//     un-contracted f32 and round-half-even; cos/sin are the glibc 2.39 cosf/sinf polynomials in f64
//     (bit-equal to the host libm the oracle was pinned against, tests/golden/sincos.json).
// No blurred pyramid is ever written to HBM.
#include "orbx_internal.cuh"

#define DESC_WARPS 8
#define PW 43            // patch width/height
#define PWORDS 12        // 32-bit words per staged patch row (48 bytes)
#define HBP 40           // pitch (u16) of the horizontally blurred patch, 37 valid columns (+3 scratch)

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

__global__ void __launch_bounds__(DESC_WARPS * 32) describe_kernel(OrbxFrameLayout L, OrbxKp28* __restrict__ kps,
                                                                    uint8_t* __restrict__ desc, int cap,
                                                                    int* __restrict__ nkp)
{
    __shared__ __align__(16) uint32_t s_raw[DESC_WARPS][PW * PWORDS + 4];   // +4: the last row's aligned windows over-read by up to two words
    __shared__ __align__(16) unsigned short s_hb[DESC_WARPS][PW * HBP];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // this lane's 16 sample points (32 int8 = two 128-bit words), fetched first so the latency hides behind staging;
    // no block-level barrier anywhere in this kernel
    const uint4 pat_lo = __ldg(reinterpret_cast<const uint4*>(g_pattern32) + 2 * lane);
    const uint4 pat_hi = __ldg(reinterpret_cast<const uint4*>(g_pattern32) + 2 * lane + 1);
    const int frame = blockIdx.y;
    const int ord = blockIdx.x * DESC_WARPS + wid;      // keypoint ordinal inside the frame (level-major)
    const int* cnt = L.lvl_kp_count + (size_t)frame * L.nlevels;
    int level = -1, k = 0, total = 0;
    for (int l = 0; l < L.nlevels; l++) {
        const int c = cnt[l];
        if (level < 0 && ord < total + c) { level = l; k = ord - total; }
        total += c;
    }
    if (ord == 0 && lane == 0) nkp[frame] = total;
    if (level < 0 || ord >= cap) return;                 // warp-uniform
    const OrbxLevelGeom g = L.lvl[level];
    const uint32_t pk = L.lvl_kp[(size_t)frame * L.kp_cap_total + L.lvl_kp_off[level] + k];
    const int kx = pk & 0xfff, ky = (pk >> 12) & 0xfff, score = pk >> 24;

    // ---- stage the 43x43 raw patch (rows ky-21.., columns kx-21..) with aligned 32-bit loads
    const uint8_t* p0 = L.raw + (size_t)frame * L.frame_raw_bytes + g.raw_off +
                        (size_t)(ky - 21 + ORBX_EDGE) * g.pitch + (kx - 21 + ORBX_XOFF);
    const int sh = (int)(reinterpret_cast<uintptr_t>(p0) & 3);
    const uint8_t* pa = p0 - sh;
    uint32_t* raw32 = s_raw[wid];
    {
        // 516 words, 32 per step; loads are issued in batches so that several are in flight per lane
        const uint32_t* pa32 = reinterpret_cast<const uint32_t*>(pa);
        const int pitch_w = g.pitch >> 2;
#pragma unroll
        for (int b0 = 0; b0 < PW * PWORDS; b0 += 32 * 6) {
            uint32_t v[6];
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const int i = b0 + 32 * u + lane;
                const int r = i / PWORDS, c = i - r * PWORDS;
                v[u] = i < PW * PWORDS ? __ldg(pa32 + r * pitch_w + c) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 6; u++) {
                const int i = b0 + 32 * u + lane;
                if (i < PW * PWORDS) raw32[i] = v[u];
            }
        }
    }
    __syncwarp();
    const uint8_t* raw8 = reinterpret_cast<const uint8_t*>(raw32) + sh;   // raw8[r*48 + c], c in [0,43)

    // ---- IC_Angle on the un-blurred level
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        // rows +v and -v together, as the reference does (ORBextractor.cc:91-102); disc half-widths are literals
        constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        const int u = lane - 15;
        const int au = u < 0 ? -u : u;
        const uint8_t* ctr = raw8 + 21 * (PWORDS * 4) + 21 + u;
        m10 = u * ctr[0];
#pragma unroll
        for (int v = 1; v <= 15; v++) {
            if (au <= UMAX[v]) {
                const int vp = ctr[v * (PWORDS * 4)], vm = ctr[-v * (PWORDS * 4)];
                m10 += u * (vp + vm);
                m01 += v * (vp - vm);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = dev_fast_atan2((float)m01, (float)m10);

    // ---- horizontal pass of the fixed-point Gaussian on the patch: hb[r][c] <-> patch column c+3
    unsigned short* hb = s_hb[wid];
    // each lane produces 8 adjacent outputs of a row from five aligned words: the twelve 4-byte windows starting at
    // bytes 0..11 come from funnel shifts, and every output is two 4-way byte dot products (weights 18,34,48,56 | 48,34,18,0)
    const uint32_t WT0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), WT1 = 48u | (34u << 8) | (18u << 16);
    for (int gi = lane; gi < PW * 5; gi += 32) {
        const int r = gi / 5, c0 = (gi - r * 5) * 8;
        const int bo = sh + c0;
        // bo >> 2 == c0 / 4 is even (sh < 4), so the five words are two aligned 64-bit loads and one 32-bit load
        const uint32_t* rw = raw32 + r * PWORDS + (bo >> 2);
        const int s8 = (bo & 3) * 8;
        const uint2 w01 = *reinterpret_cast<const uint2*>(rw), w23 = *reinterpret_cast<const uint2*>(rw + 2);
        const uint32_t w0 = w01.x, w1 = w01.y, w2 = w23.x, w3 = w23.y, w4 = rw[4];
        uint32_t W[12];
        W[0] = __funnelshift_r(w0, w1, s8); W[4] = __funnelshift_r(w1, w2, s8);
        W[8] = __funnelshift_r(w2, w3, s8);
        const uint32_t a3 = __funnelshift_r(w3, w4, s8);
#pragma unroll
        for (int t = 1; t < 4; t++) {
            W[t] = __funnelshift_r(W[0], W[4], 8 * t);
            W[4 + t] = __funnelshift_r(W[4], W[8], 8 * t);
            W[8 + t] = __funnelshift_r(W[8], a3, 8 * t);
        }
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = __dp4a(W[j + 4], WT1, __dp4a(W[j], WT0, 0u));
        // outputs are < 2^16: two per 32-bit word, one 128-bit store (rows are 80 bytes, groups 16 bytes apart)
        *reinterpret_cast<uint4*>(hb + r * HBP + c0) = make_uint4(o[0] | (o[1] << 16), o[2] | (o[3] << 16), o[4] | (o[5] << 16), o[6] | (o[7] << 16));
    }
    __syncwarp();

    // ---- steered rBRIEF: lane i -> descriptor byte i
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float a, b;
    dev_sincosf(__fmul_rn(angle, factorPI), &b, &a);
    const uint32_t patw[8] = {pat_lo.x, pat_lo.y, pat_lo.z, pat_lo.w, pat_hi.x, pat_hi.y, pat_hi.z, pat_hi.w};
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const uint32_t w = patw[t];                    // (x0, y0, x1, y1) of test t as four int8
        int smp[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const float px = (float)(signed char)(w >> (16 * e)), py = (float)(signed char)(w >> (16 * e + 8));
            const int iy = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int ix = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            const unsigned short* h = hb + (iy + 18) * HBP + (ix + 18);
            const unsigned acc = 18u * (h[0] + h[6 * HBP]) + 34u * (h[HBP] + h[5 * HBP]) + 48u * (h[2 * HBP] + h[4 * HBP]) + 56u * h[3 * HBP];
            smp[e] = (int)((acc + 32768u) >> 16);
        }
        val |= (smp[0] < smp[1]) << t;
    }
    const size_t o = (size_t)frame * cap + ord;
    desc[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        OrbxKp28 kp;
        kp.x = level ? __fmul_rn((float)kx, g.scale) : (float)kx;
        kp.y = level ? __fmul_rn((float)ky, g.scale) : (float)ky;
        kp.size = g.kp_size;
        kp.angle = angle;
        kp.response = (float)score;
        kp.octave = level;
        kp.class_id = -1;
        kps[o] = kp;
    }
}

void orbx_launch_describe(const OrbxFrameLayout& L, int nframes, OrbxKp28* d_kps, uint8_t* d_desc, int cap, int* d_nkp,
                          cudaStream_t st)
{
    int total_cap = L.kp_cap_total;
    dim3 grid((total_cap + DESC_WARPS - 1) / DESC_WARPS, nframes);
    describe_kernel<<<grid, DESC_WARPS * 32, 0, st>>>(L, d_kps, d_desc, cap, d_nkp);
}
