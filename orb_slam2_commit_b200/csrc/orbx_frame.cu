// orbx_frame.cu — per-keypoint work that Frame's constructors do right after extraction (Frame.cc:86-110):
//   * Frame::UndistortKeyPoints / Frame::ComputeImageBounds (Frame.cc:471-538) = cv::undistortPoints(mat, mat, mK,
//     mDistCoef, cv::Mat(), mK): OpenCV 4.x iterates x <- (x0 - delta(x)) / cdist(x) five times in DOUBLE and projects
//     back with K. Reproduced operation for operation with un-contracted f64 (no FMA), so mvKeysUn is bit-identical.
#include "orbx_internal.cuh"

// one thread per keypoint; kp records are 28 bytes (x, y first). K and the distortion come in as doubles converted on
// the host exactly like cv::Mat::convertTo(CV_64F) does (f32 -> f64 is exact).
__global__ void __launch_bounds__(256) undistort_kernel(const OrbxKp28* __restrict__ in, OrbxKp28* __restrict__ out, int n,
                                                        OrbxUndistortArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    OrbxKp28 kp = in[i];
    const double u = (double)kp.x, v = (double)kp.y;
    double x = __dmul_rn(__dsub_rn(u, a.cx), a.ifx), y = __dmul_rn(__dsub_rn(v, a.cy), a.ify);
    const double x0 = x, y0 = y;
    const double k0 = a.k[0], k1 = a.k[1], p1 = a.k[2], p2 = a.k[3], k4 = a.k[4];
    const double tp1 = __dmul_rn(2.0, p1), tp2 = __dmul_rn(2.0, p2);
#pragma unroll 1
    for (int j = 0; j < 5; j++) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        // icdist = 1 / (1 + ((k3 r2 + k2) r2 + k1) r2); the rational numerator (k4..k6) is absent in ORB-SLAM2's models
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k4, r2), k1), r2), k0), r2));
        const double icdist = __ddiv_rn(1.0, den);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(tp1, x), y), __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), __dmul_rn(__dmul_rn(tp2, x), y));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    kp.x = __double2float_rn(__dadd_rn(__dmul_rn(a.fx, x), a.cx));
    kp.y = __double2float_rn(__dadd_rn(__dmul_rn(a.fy, y), a.cy));
    out[i] = kp;
}

void orbx_launch_undistort(const OrbxKp28* d_in, OrbxKp28* d_out, int n, const OrbxUndistortArgs& a, cudaStream_t st)
{
    if (n <= 0) return;
    undistort_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_in, d_out, n, a);
}

// cv::initUndistortRectifyMap(K, D, R, P, size, CV_32F, M1, M2) (Examples/Stereo/stereo_euroc.cc:96-97): one thread per map
// pixel, OpenCV 4.x's double arithmetic operation for operation with un-contracted f64 — (_x, _y, _w) = iR * (j, i, 1) as
// (i*ir[1] + ir[2]) + j*ir[0], the Brown-Conrady model with k1..k6, p1, p2, s1..s4, projection with K, rounded to f32 once.
// Writes the f32 map pair and / or the fixed-point form pyr_level0_remap_kernel reads (cvRound(map * 32), OpenCV's INTER_BITS = 5
// conversion), so a camera goes from its calibration to a rectifying extractor without a host round trip.
__global__ void __launch_bounds__(256) rectify_map_kernel(OrbxRectifyArgs a, int w, int h, float* __restrict__ map1,
                                                          float* __restrict__ map2, uint2* __restrict__ fixed)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= w || i >= h) return;
    const double di = (double)i, dj = (double)j;
    const double _x = __dadd_rn(__dadd_rn(__dmul_rn(di, a.ir[1]), a.ir[2]), __dmul_rn(dj, a.ir[0]));
    const double _y = __dadd_rn(__dadd_rn(__dmul_rn(di, a.ir[4]), a.ir[5]), __dmul_rn(dj, a.ir[3]));
    const double _w = __dadd_rn(__dadd_rn(__dmul_rn(di, a.ir[7]), a.ir[8]), __dmul_rn(dj, a.ir[6]));
    const double iw = __ddiv_rn(1.0, _w), x = __dmul_rn(_x, iw), y = __dmul_rn(_y, iw);
    const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y), r2 = __dadd_rn(x2, y2), _2xy = __dmul_rn(__dmul_rn(2.0, x), y);
    const double k1 = a.k[0], k2 = a.k[1], p1 = a.k[2], p2 = a.k[3], k3 = a.k[4], k4 = a.k[5], k5 = a.k[6], k6 = a.k[7];
    const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k3, r2), k2), r2), k1), r2));
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k6, r2), k5), r2), k4), r2));
    const double kr = __ddiv_rn(num, den);
    // xd = x*kr + p1*_2xy + p2*(r2 + 2*x2) + s1*r2 + s2*r2*r2, summed left to right
    double xd = __dadd_rn(__dmul_rn(x, kr), __dmul_rn(p1, _2xy));
    xd = __dadd_rn(xd, __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    xd = __dadd_rn(xd, __dmul_rn(a.k[8], r2));
    xd = __dadd_rn(xd, __dmul_rn(__dmul_rn(a.k[9], r2), r2));
    double yd = __dadd_rn(__dmul_rn(y, kr), __dmul_rn(p1, __dadd_rn(r2, __dmul_rn(2.0, y2))));
    yd = __dadd_rn(yd, __dmul_rn(p2, _2xy));
    yd = __dadd_rn(yd, __dmul_rn(a.k[10], r2));
    yd = __dadd_rn(yd, __dmul_rn(__dmul_rn(a.k[11], r2), r2));
    const float u = __double2float_rn(__dadd_rn(__dmul_rn(a.fx, xd), a.u0));
    const float v = __double2float_rn(__dadd_rn(__dmul_rn(a.fy, yd), a.v0));
    const size_t o = (size_t)i * w + j;
    if (map1) map1[o] = u;
    if (map2) map2[o] = v;
    if (fixed) {
        const int sx = __float2int_rn(__fmul_rn(u, 32.0f)), sy = __float2int_rn(__fmul_rn(v, 32.0f));
        const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
        fixed[o] = make_uint2(((uint32_t)ix & 0xffffu) | ((uint32_t)iy << 16), (uint32_t)(sx & 31) | ((uint32_t)(sy & 31) << 5));
    }
}

void orbx_launch_rectify_map(const OrbxRectifyArgs& a, int w, int h, float* d_map1, float* d_map2, uint2* d_fixed, cudaStream_t st)
{
    if (w <= 0 || h <= 0) return;
    rectify_map_kernel<<<dim3((w + 255) / 256, h), 256, 0, st>>>(a, w, h, d_map1, d_map2, d_fixed);
}
