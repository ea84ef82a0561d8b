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
