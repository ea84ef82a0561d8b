// orbx_describe.cu — orientation + 7x7 Gaussian + steered rBRIEF, one warp per keypoint (replaces
// computeOrientation/IC_Angle ORBextractor.cc:77-105,492-499, the per-level cv::GaussianBlur :1188-1190 and
// computeOrbDescriptor :110-152, plus the coordinate scaling / output packing of operator() :1194-1209).
//
// One elected lane has the TMA unit copy the 64x43-byte box around the keypoint from the HBM pyramid into the warp's
// shared-memory slot (cp.async.bulk.tensor through the level's tensor map, completion on the warp's mbarrier); then
//   * IC_Angle: lane = column u in [-15,15], integer moments, warp-shuffle reduce, cv::fastAtan2 polynomial
//     evaluated with un-contracted f32 mul/add (bit-equal to OpenCV's scalar path);
//   * blur of the patch only: the reference blurs the whole level and then reads 512 points per keypoint; here both passes
//     of OpenCV's fixed-point kernel [18,34,48,56,48,34,18]/256 run on the 43x43 patch in shared memory (vertical pass
//     exact in 16 bits, horizontal pass ((sum + 2^15) >> 16) to a 37x37 byte image that overwrites the raw patch), so a
//     sample is ONE byte load — the kernel is bound by shared-memory wavefronts, and the on-demand form of the second pass
//     (four random 32-bit loads per sample, 64 per lane) was more than half of them. Border handling is the level's own
//     REFLECT_101, which is exactly what the 19-px apron in HBM holds;
//   * rBRIEF: lane i builds descriptor byte i (8 tests); sample = center + cvRound(x*b+y*a, x*a-y*b) with
//     un-contracted f32 and round-half-even; cos/sin are the glibc 2.39 cosf/sinf polynomials in f64
//     (bit-equal to the host libm the oracle was pinned against, tests/golden/sincos.json).
// No blurred pyramid is ever written to HBM.
#include "orbx_internal.cuh"
#include "orbx_tma.cuh"
#include <type_traits>

#ifndef DESC_WARPS
#define DESC_WARPS 4         // warps (= keypoints) per CTA
#endif
#define PW 43            // patch width/height
#define PWORDS 12        // 32-bit words of a staged patch row that the kernel works on (48 bytes: 43 + up to 3 of alignment)
#define RPW 16           // pitch of the staged rows in words: the 64-byte TMA box (its origin must be 16-byte aligned)
#define VROWS 37         // rows of the vertically blurred patch (patch rows 3..39 centred)
#define VPW 25           // its pitch in 32-bit words: 48 u16 per row, one per byte of the staged 48-byte patch row, + 1 word: an odd
                         // pitch puts the 32 rows the lanes of a step read in the second pass into 32 different banks
#define BP 40            // pitch of the blurred 37x37 byte image (8 outputs per lane and step, 5 steps per row)

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

#define RAW_SLOT_WORDS 704   // 64 x 43 bytes = 688 words, padded to a multiple of 128 bytes (TMA destination alignment)
__global__ void __launch_bounds__(DESC_WARPS * 32) describe_kernel(OrbxFrameLayout L, const __grid_constant__ OrbxTmaps maps,
                                                                    OrbxKp28* __restrict__ kps, uint8_t* __restrict__ desc, int cap,
                                                                    int* __restrict__ nkp)
{
    __shared__ __align__(128) uint32_t s_raw[DESC_WARPS][RAW_SLOT_WORDS];
    __shared__ __align__(16) uint32_t s_vb[DESC_WARPS][VROWS * VPW + 8];   // + 8: the last row's windows over-read by up to four words
    __shared__ __align__(8) unsigned long long s_bar[DESC_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // this lane's 16 sample points (32 int8 = two 128-bit words), fetched first so the latency hides behind staging;
    // no block-level barrier anywhere in this kernel
    const uint4 pat_lo = __ldg(reinterpret_cast<const uint4*>(g_pattern32) + 2 * lane);
    const uint4 pat_hi = __ldg(reinterpret_cast<const uint4*>(g_pattern32) + 2 * lane + 1);
    const int frame = blockIdx.y;
    const int ord = blockIdx.x * DESC_WARPS + wid;      // keypoint ordinal inside the frame (level-major)
    const int* cnt = L.lvl_kp_count + (size_t)frame * L.nlevels;
    // level of the ordinal: lane l holds the count of level l, a warp scan gives the running totals (nlevels <= 16)
    int level = -1, k = 0, total;
    {
        const int c = lane < L.nlevels ? cnt[lane] : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < ORBX_MAX_LEVELS; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ord < incl);
        total = __shfl_sync(0xffffffffu, incl, ORBX_MAX_LEVELS - 1);   // the scan spans 16 lanes
        if (m) { level = __ffs(m) - 1; k = ord - __shfl_sync(0xffffffffu, incl - c, level); }
    }
    if (ord == 0 && lane == 0) nkp[frame] = total;
    if (level < 0 || ord >= cap) return;                 // warp-uniform
    const OrbxLevelGeom g = L.lvl[level];
    const uint32_t pk = L.lvl_kp[(size_t)frame * L.kp_cap_total + L.lvl_kp_off[level] + k];
    const int kx = pk & 0xfff, ky = (pk >> 12) & 0xfff, score = pk >> 24;

    // ---- stage the raw patch: rows ky-21 .. ky+21 of the 64-byte box that starts at the 16-byte aligned column at or left
    // of kx-21; the kernel then works on the 48-byte window that starts at the WORD holding column kx-21 (byte `sh` of it)
    uint32_t* slot32 = s_raw[wid];
    const int x0 = kx - 21 + ORBX_XOFF;
    {
        const uint32_t bar = orbx_smem_addr(&s_bar[wid]);
        if (lane == 0) {
            orbx_mbar_init(bar, 1);
            orbx_mbar_expect_tx(bar, PW * RPW * 4);
            orbx_tma_load_3d(orbx_smem_addr(slot32), &maps.m[level], x0 & ~15, ky - 21 + ORBX_EDGE, L.frame0 + frame, bar);
        }
        __syncwarp();                                    // the barrier is initialised before anyone waits on it
        orbx_mbar_wait(bar, 0);
    }
    const uint32_t* raw32 = slot32 + ((x0 & 15) >> 2);   // row r of the window = raw32 + r * RPW, 12 words
    const int sh = x0 & 3;
    const uint8_t* raw8 = reinterpret_cast<const uint8_t*>(raw32) + sh;   // raw8[r*64 + c], c in [0,43)

    // ---- IC_Angle on the un-blurred level
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        // rows +v and -v together, as the reference does (ORBextractor.cc:91-102); disc half-widths are literals
        constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        const int u = lane - 15;
        const int au = u < 0 ? -u : u;
        const uint8_t* ctr = raw8 + 21 * (RPW * 4) + 21 + u;
        m10 = u * ctr[0];
#pragma unroll
        for (int v = 1; v <= 15; v++) {
            if (au <= UMAX[v]) {
                const int vp = ctr[v * (RPW * 4)], vm = ctr[-v * (RPW * 4)];
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

    // ---- first pass of the fixed-point Gaussian on the patch, taken VERTICALLY: vb[ro][b] = sum_k w_k * row(ro+k)[b]
    // for every byte position b of the staged 48-byte rows (exact in 16 bits). OpenCV runs the horizontal pass first,
    // but neither pass rounds before the final (sum + 2^15) >> 16, so the order does not change the result; with the
    // vertical pass first the seven taps a sample point needs afterwards are CONTIGUOUS in shared memory: four 32-bit
    // loads instead of seven 16-bit ones (this kernel is bound by shared-memory wavefronts: sample reads hit random banks).
    // Lane = (8-byte column) x (strip of 8 output rows): a 7-row window of byte pairs spread into 16-bit halves slides
    // down the column, two outputs per 32-bit op (every sum stays below 2^16).
    uint32_t* vb = s_vb[wid];
    if (lane < 30) {
        const int col = lane % 6, strip = lane / 6;
        // strips start at rows 0, 7, 14, 21, 29 (eight rows each, three rows computed twice): with the 16-word pitch of the
        // staged rows, odd and even starts sit 16 banks apart, which takes the loads from 5-way to 3-way conflicts
        const int ro0 = strip < 4 ? strip * 7 : VROWS - 8;
        // (the window starts at a word, not at an 8-byte boundary of the box: two 32-bit loads per row)
        const uint32_t* src = raw32 + ro0 * RPW + 2 * col;
        uint32_t E[7][4];
        auto expand = [&](uint32_t (&e)[4], const uint32_t wx, const uint32_t wy) {
            e[0] = __byte_perm(wx, 0, 0x4140); e[1] = __byte_perm(wx, 0, 0x4342);
            e[2] = __byte_perm(wy, 0, 0x4140); e[3] = __byte_perm(wy, 0, 0x4342);
        };
#pragma unroll
        for (int i = 0; i < 6; i++) expand(E[i], src[i * RPW], src[i * RPW + 1]);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            expand(E[(j + 6) % 7], src[(j + 6) * RPW], src[(j + 6) * RPW + 1]);
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                o[q] = 18u * (E[j % 7][q] + E[(j + 6) % 7][q]) + 34u * (E[(j + 1) % 7][q] + E[(j + 5) % 7][q]) +
                       48u * (E[(j + 2) % 7][q] + E[(j + 4) % 7][q]) + 56u * E[(j + 3) % 7][q];
            uint32_t* vo = vb + (ro0 + j) * VPW + col * 4;
            vo[0] = o[0]; vo[1] = o[1]; vo[2] = o[2]; vo[3] = o[3];
        }
    }
    __syncwarp();

    // ---- second (horizontal) pass, dense: blur[r][c] = (sum_t w_t * vb[r][sh + c + t] + 2^15) >> 16 for c in [0, 37); a lane
    // takes 8 adjacent outputs of one row per step (185 units = 37 rows x 5 groups over 32 lanes), i.e. 14 contiguous u16 =
    // eight word loads, four dp2a per output with the weight pairs laid out for the parity of its first tap (the parity of
    // sh decides between the two instances, warp-uniform). Units run column-major, so a step's lanes read 32 rows of the odd-pitch
    // vb at one column offset: conflict-free. The byte image overwrites the raw patch, which nobody reads
    // any more (IC_Angle and the vertical pass are done).
    uint8_t* blur = reinterpret_cast<uint8_t*>(slot32);
    {
        const uint32_t* vrow = vb + (sh >> 1);
        auto hpass = [&](auto PARC) {
            constexpr int PAR = decltype(PARC)::value;
            for (int u = lane; u < VROWS * 5; u += 32) {
                const int g8 = u / VROWS, r = u - VROWS * g8;      // column-major: the lanes of a step take (mostly) 32 consecutive rows
                const uint32_t* h = vrow + r * VPW + 4 * g8;
                uint32_t W[8];
#pragma unroll
                for (int i = 0; i < 7 + PAR; i++) W[i] = h[i];
                if (PAR == 0) W[7] = 0u;
                uint32_t acc[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int m = (PAR + j) >> 1;
                    if ((PAR + j) & 1) {
                        unsigned a_ = __dp2a_lo(W[m], 18u << 8, 32768u);
                        a_ = __dp2a_lo(W[m + 1], 34u | (48u << 8), a_);
                        a_ = __dp2a_lo(W[m + 2], 56u | (48u << 8), a_);
                        acc[j] = __dp2a_lo(W[m + 3], 34u | (18u << 8), a_);
                    } else {
                        unsigned a_ = __dp2a_lo(W[m], 18u | (34u << 8), 32768u);
                        a_ = __dp2a_lo(W[m + 1], 48u | (56u << 8), a_);
                        a_ = __dp2a_lo(W[m + 2], 48u | (34u << 8), a_);
                        acc[j] = __dp2a_lo(W[m + 3], 18u, a_);
                    }
                }
                // byte 2 of every accumulator is the blurred pixel ((sum + 2^15) >> 16 < 256)
                const uint32_t lo = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
                const uint32_t hi = __byte_perm(__byte_perm(acc[4], acc[5], 0x0062), __byte_perm(acc[6], acc[7], 0x0062), 0x5410);
                *reinterpret_cast<uint2*>(blur + r * BP + 8 * g8) = make_uint2(lo, hi);
            }
        };
        __syncwarp();
        if (sh & 1) hpass(std::integral_constant<int, 1>{}); else hpass(std::integral_constant<int, 0>{});
    }
    __syncwarp();

    // ---- steered rBRIEF: lane i -> descriptor byte i
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.0);
    float a, b;
    dev_sincosf(__fmul_rn(angle, factorPI), &b, &a);
    const uint32_t patw[8] = {pat_lo.x, pat_lo.y, pat_lo.z, pat_lo.w, pat_hi.x, pat_hi.y, pat_hi.z, pat_hi.w};
    const uint8_t* bc = blur + 18 * BP + 18;            // blurred pixel at the keypoint
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
            smp[e] = bc[iy * BP + ix];
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

void orbx_launch_describe(const OrbxFrameLayout& L, const OrbxTmaps& maps, int nframes, OrbxKp28* d_kps, uint8_t* d_desc, int cap,
                          int* d_nkp, cudaStream_t st)
{
    int total_cap = L.kp_cap_total;
    dim3 grid((total_cap + DESC_WARPS - 1) / DESC_WARPS, nframes);
    describe_kernel<<<grid, DESC_WARPS * 32, 0, st>>>(L, maps, d_kps, d_desc, cap, d_nkp);
}
