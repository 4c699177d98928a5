// K11 / K16: device-side geometry for the camera-ray head and the depth window aligner.
//
//   l4p_pose_from_rays   Plücker ray map -> camera centres (3x3 normal equations), optional fixed intrinsics
//                        from frame 0 (normalised DLT homography + consensus refits + RQ), per-frame Kabsch
//                        rotation (3x3 one-sided Jacobi SVD), extrinsics and their inverse (pose).
//                        Replaces geometry_utils.py:249-282,285-305,308-328,331-406,409-456,493-579 and the
//                        per-(b,t) Python SVD loop / cv2 host round trip in them.
//   l4p_affine_align_*   5-moment reduction + closed-form 2x2 solve + apply for LstSqAffineAligner
//                        (aligner.py:45-66, misc.py:48-62) (K16).
//
// All small dense algebra is done in fp64 by one thread per problem; reductions over rays are block-wide.
#include "common.cuh"

namespace l4p {

// ---------------------------------------------------------------------------------------------- helpers
template <int N>
L4P_DEVICE void block_sum(double* v, double* smem /* [N * 32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) smem[i * 32 + warp] = x;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += smem[i * 32 + w];
    v[i] = s;
  }
  __syncthreads();
}

__device__ double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ void inv3(const double* m, double* r) {
  const double d = det3(m);
  const double id = 1.0 / d;
  r[0] = (m[4] * m[8] - m[5] * m[7]) * id; r[1] = (m[2] * m[7] - m[1] * m[8]) * id; r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  r[3] = (m[5] * m[6] - m[3] * m[8]) * id; r[4] = (m[0] * m[8] - m[2] * m[6]) * id; r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  r[6] = (m[3] * m[7] - m[4] * m[6]) * id; r[7] = (m[1] * m[6] - m[0] * m[7]) * id; r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
__device__ void mul3(const double* a, const double* b, double* c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

// Eigenvector of the SMALLEST eigenvalue of a symmetric positive semi-definite 9x9 matrix (the DLT normal matrix) by shifted
// inverse iteration on a Cholesky factor. Every loop has compile-time bounds, so the factor lives in registers; round 1 ran
// a cyclic Jacobi here (81 + 81 doubles in local memory, 6-9 sweeps x 36 rotations: 0.24 ms per solve on one thread, 4 solves
// per call = 3 % of the whole all-heads step). Convergence factor (l1 + mu) / (l2 + mu): 1-2 iterations
// on consistent correspondences (l1 ~ 0); near-degenerate problems (l1 ~ l2: any vector of that eigenspace is as good as
// another) stop at the iteration cap (24). Returns the Rayleigh quotient.
__device__ double smallest_eigvec9(const double* M, double* x) {
  constexpr int n = 9;
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < n; ++i) tr += M[i * n + i];
  double mu = 1e-13 * tr + 1e-300;
  double L[n * (n + 1) / 2], rd[n];
  for (int attempt = 0; attempt < 4; ++attempt) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double a = M[i * n + j] + (i == j ? mu : 0.0);
#pragma unroll
        for (int k = 0; k < j; ++k) a -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
        if (i == j) {
          if (!(a > 0.0)) { ok = false; a = 1.0; }
          L[i * (i + 1) / 2 + j] = sqrt(a);
          rd[i] = 1.0 / L[i * (i + 1) / 2 + j];      // reciprocal diagonal: the solves below multiply instead of dividing
        } else {
          L[i * (i + 1) / 2 + j] = a * rd[j];
        }
      }
    }
    if (ok) break;
    mu = mu * 1e3 + 1e-12 * tr;   // round-off made a pivot non-positive: shift harder
  }
  double v[n];
#pragma unroll
  for (int i = 0; i < n; ++i) v[i] = 1.0 / 3.0 + 0.01 * i;   // generic start (not orthogonal to any eigenvector by symmetry)
  for (int it = 0; it < 24; ++it) {
    double y[n];
#pragma unroll
    for (int i = 0; i < n; ++i) {          // L y = v
      double a = v[i];
#pragma unroll
      for (int k = 0; k < i; ++k) a -= L[i * (i + 1) / 2 + k] * y[k];
      y[i] = a * rd[i];
    }
#pragma unroll
    for (int i = n - 1; i >= 0; --i) {     // L^T z = y  (z overwrites y)
      double a = y[i];
#pragma unroll
      for (int k = i + 1; k < n; ++k) a -= L[k * (k + 1) / 2 + i] * y[k];
      y[i] = a * rd[i];
    }
    double nn = 0.0, dot = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) nn += y[i] * y[i];
    const double inv = 1.0 / sqrt(nn);
#pragma unroll
    for (int i = 0; i < n; ++i) { y[i] *= inv; dot += y[i] * v[i]; }
    const double sgn = dot < 0.0 ? -1.0 : 1.0;
    double diff = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) { const double d = sgn * y[i] - v[i]; diff += d * d; v[i] = sgn * y[i]; }
    if (diff < 1e-28 && it > 0) break;
  }
  double lam = 0.0;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) a += M[i * n + j] * v[j];
    lam += a * v[i];
    x[i] = v[i];
  }
  return lam;
}

// 3x3 SVD by one-sided Jacobi: H = U diag(s) V^T, singular values sorted descending.
__device__ void svd3(const double* H, double* U, double* S, double* V) {
  double W[9];
  for (int i = 0; i < 9; ++i) W[i] = H[i];
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double rot = 0.0;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) {
          alpha += W[k * 3 + p] * W[k * 3 + p];
          beta += W[k * 3 + q] * W[k * 3 + q];
          gamma += W[k * 3 + p] * W[k * 3 + q];
        }
        if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 1e-15 * sqrt(alpha * beta)) continue;
        rot += fabs(gamma);
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          const double wp = W[k * 3 + p], wq = W[k * 3 + q];
          W[k * 3 + p] = c * wp - s * wq;
          W[k * 3 + q] = s * wp + c * wq;
          const double vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - s * vq;
          V[k * 3 + q] = s * vp + c * vq;
        }
      }
    if (rot == 0.0) break;
  }
  for (int j = 0; j < 3; ++j) S[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
  // sort descending
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2 - a; ++b)
      if (S[b] < S[b + 1]) {
        double t = S[b]; S[b] = S[b + 1]; S[b + 1] = t;
        for (int k = 0; k < 3; ++k) {
          t = W[k * 3 + b]; W[k * 3 + b] = W[k * 3 + b + 1]; W[k * 3 + b + 1] = t;
          t = V[k * 3 + b]; V[k * 3 + b] = V[k * 3 + b + 1]; V[k * 3 + b + 1] = t;
        }
      }
  for (int j = 0; j < 3; ++j) {
    if (S[j] > 1e-12 * (S[0] + 1e-300)) {
      for (int k = 0; k < 3; ++k) U[k * 3 + j] = W[k * 3 + j] / S[j];
    } else {
      for (int k = 0; k < 3; ++k) U[k * 3 + j] = 0.0;
    }
  }
  // complete U to an orthonormal basis if rank deficient (third column = u0 x u1)
  if (!(S[2] > 1e-12 * (S[0] + 1e-300))) {
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
}

// geometry_utils.py:119-125 on a 3x3 block (row-major K)
__device__ void denorm_k(double* k, int h, int w) {
  for (int j = 0; j < 3; ++j) { k[j] *= w; k[3 + j] *= h; }
  k[2] -= 0.5; k[5] -= 0.5;
}
__device__ void norm_k(double* k, int h, int w) {
  k[2] += 0.5; k[5] += 0.5;
  for (int j = 0; j < 3; ++j) { k[j] /= w; k[3 + j] /= h; }
}

// ---------------------------------------------------------------------------------------------- intrinsics (frame 0)
// One block per batch element. kgrid [B,9] (ray-grid units, used for the ideal rays), kout [B,4,4,T] fp32.
__global__ void __launch_bounds__(256)
estimate_k_kernel(const float* __restrict__ rays, int T, int h, int w, int outH, int outW, float thr, int refits,
                  double* __restrict__ kgrid, float* __restrict__ kout) {
  __shared__ double red[45 * 32];
  __shared__ double sh_A[9];
  __shared__ double sh_norm[6];
  extern __shared__ unsigned char inl[];  // [h*w] inlier flags
  const int b = blockIdx.x;
  const int n = h * w;
  const long long plane = (long long)T * n;
  const float* base = rays + (long long)b * 6 * plane;  // frame 0: offset 0 inside each channel plane

  for (int it = 0; it <= refits; ++it) {
    // --- normalisation statistics over the active set
    double st[5] = {0, 0, 0, 0, 0};  // sum ox, oy, tx, ty, count
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float dz = base[2 * plane + i];
      const bool valid = fabsf(dz) > 1e-4f;  // origin z = 1/|(i,j,1)| is always > 1e-4 on these grids
      bool use = valid;
      if (it > 0) use = use && inl[i];
      if (it == 0) inl[i] = valid;
      if (use) {
        st[0] += (double)(i % w); st[1] += (double)(i / w);
        st[2] += (double)base[i] / dz; st[3] += (double)base[plane + i] / dz;
        st[4] += 1.0;
      }
    }
    block_sum<5>(st, red);
    const double cnt = st[4] > 0 ? st[4] : 1.0;
    const double mox = st[0] / cnt, moy = st[1] / cnt, mtx = st[2] / cnt, mty = st[3] / cnt;
    double sd[2] = {0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (inl[i]) {
        const float dz = base[2 * plane + i];
        const double ox = (double)(i % w) - mox, oy = (double)(i / w) - moy;
        const double tx = (double)base[i] / dz - mtx, ty = (double)base[plane + i] / dz - mty;
        sd[0] += sqrt(ox * ox + oy * oy);
        sd[1] += sqrt(tx * tx + ty * ty);
      }
    }
    block_sum<2>(sd, red);
    const double so = sd[0] > 0 ? 1.4142135623730951 * cnt / sd[0] : 1.0;
    const double stt = sd[1] > 0 ? 1.4142135623730951 * cnt / sd[1] : 1.0;
    // --- A^T A of the DLT system (45 unique entries)
    double acc[45];
    for (int i = 0; i < 45; ++i) acc[i] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (inl[i]) {
        const float dz = base[2 * plane + i];
        const double x = ((double)(i % w) - mox) * so, y = ((double)(i / w) - moy) * so;
        const double u = ((double)base[i] / dz - mtx) * stt, v = ((double)base[plane + i] / dz - mty) * stt;
        const double r1[9] = {-x, -y, -1, 0, 0, 0, u * x, u * y, u};
        const double r2[9] = {0, 0, 0, -x, -y, -1, v * x, v * y, v};
        int k = 0;
        for (int p = 0; p < 9; ++p)
          for (int q = p; q < 9; ++q) acc[k++] += r1[p] * r1[q] + r2[p] * r2[q];
      }
    }
    block_sum<45>(acc, red);
    if (threadIdx.x == 0) {
      double M[81];
      int k = 0;
#pragma unroll
      for (int p = 0; p < 9; ++p)
#pragma unroll
        for (int q = p; q < 9; ++q) { M[p * 9 + q] = acc[k]; M[q * 9 + p] = acc[k]; ++k; }
      double Hn[9];
      smallest_eigvec9(M, Hn);
      // denormalise: A = Tt^-1 Hn To,  To = [so 0 -so*mox; 0 so -so*moy; 0 0 1], Tt likewise
      const double To[9] = {so, 0, -so * mox, 0, so, -so * moy, 0, 0, 1};
      const double Tti[9] = {1.0 / stt, 0, mtx, 0, 1.0 / stt, mty, 0, 0, 1};
      double tmp[9], A[9];
      mul3(Hn, To, tmp);
      mul3(Tti, tmp, A);
      for (int i = 0; i < 9; ++i) sh_A[i] = A[i];
    }
    __syncthreads();
    if (it < refits) {
      // consensus refit: keep correspondences whose reprojection error is below the threshold
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float dz = base[2 * plane + i];
        const bool valid = fabsf(dz) > 1e-4f;
        bool ok = false;
        if (valid) {
          const double x = (double)(i % w), y = (double)(i / w);
          const double pz = sh_A[6] * x + sh_A[7] * y + sh_A[8];
          const double px = (sh_A[0] * x + sh_A[1] * y + sh_A[2]) / pz, py = (sh_A[3] * x + sh_A[4] * y + sh_A[5]) / pz;
          const double ex = px - (double)base[i] / dz, ey = py - (double)base[plane + i] / dz;
          ok = (ex * ex + ey * ey) < (double)thr * thr;
        }
        inl[i] = ok;
      }
      __syncthreads();
      // keep at least a well-posed problem: if fewer than 8 inliers survive, fall back to all valid points
      double c2[1] = {0};
      for (int i = threadIdx.x; i < n; i += blockDim.x) c2[0] += inl[i] ? 1.0 : 0.0;
      block_sum<1>(c2, red);
      if (c2[0] < 8.0) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) inl[i] = fabsf(base[2 * plane + i]) > 1e-4f;
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) {
    double A[9];
    for (int i = 0; i < 9; ++i) A[i] = sh_A[i];
    if (det3(A) < 0)
      for (int i = 0; i < 9; ++i) A[i] = -A[i];
    double Hm[9];
    inv3(A, Hm);  // H = K R
    // RQ with positive diagonal (Gram-Schmidt on the rows, bottom up) == cv2.RQDecomp3x3 up to its sign convention
    double K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double r3[3], r2[3], r1[3];
    double n3 = sqrt(Hm[6] * Hm[6] + Hm[7] * Hm[7] + Hm[8] * Hm[8]);
    for (int j = 0; j < 3; ++j) r3[j] = Hm[6 + j] / n3;
    K[8] = n3;
    K[5] = Hm[3] * r3[0] + Hm[4] * r3[1] + Hm[5] * r3[2];
    for (int j = 0; j < 3; ++j) r2[j] = Hm[3 + j] - K[5] * r3[j];
    double n2 = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    for (int j = 0; j < 3; ++j) r2[j] /= n2;
    K[4] = n2;
    K[2] = Hm[0] * r3[0] + Hm[1] * r3[1] + Hm[2] * r3[2];
    K[1] = Hm[0] * r2[0] + Hm[1] * r2[1] + Hm[2] * r2[2];
    for (int j = 0; j < 3; ++j) r1[j] = Hm[j] - K[2] * r3[j] - K[1] * r2[j];
    K[0] = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    for (int i = 0; i < 9; ++i) K[i] /= n3;
    for (int i = 0; i < 9; ++i) kgrid[b * 9 + i] = K[i];
    // report at the output resolution: denormalize(normalize(K, h, w), outH, outW)  (geometry_utils.py:575-577)
    double Ko[9];
    for (int i = 0; i < 9; ++i) Ko[i] = K[i];
    norm_k(Ko, h, w);
    denorm_k(Ko, outH, outW);
    for (int t = 0; t < T; ++t) {
      float* o = kout + (long long)b * 16 * T;
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          double val = 0.0;
          if (i < 3 && j < 3) val = Ko[i * 3 + j];
          else if (i == 3 && j == 3) val = 1.0;
          o[(i * 4 + j) * T + t] = (float)val;
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------- pose per frame
// One block per (b,t). mode 0: intrinsics from k_norm [B,4,4,T] (normalised); mode 1: kgrid [B,9].
__global__ void __launch_bounds__(256)
pose_kernel(const float* __restrict__ rays, const float* __restrict__ k_norm, const double* __restrict__ kgrid, int mode,
            int T, int h, int w, float* __restrict__ ext, float* __restrict__ pose, float* __restrict__ centers) {
  __shared__ double red[21 * 32];
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const int n = h * w;
  const long long plane = (long long)T * n;
  const float* base = rays + (long long)b * 6 * plane + (long long)t * n;

  double K[9];
  if (mode == 0) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) K[i * 3 + j] = (double)k_norm[((long long)b * 16 + i * 4 + j) * T + t];
    denorm_k(K, h, w);
  } else {
    for (int i = 0; i < 9; ++i) K[i] = kgrid[b * 9 + i];
  }
  double Ki[9];
  inv3(K, Ki);

  // sums: M = sum(I - dd^T) [6 unique], rhs = sum (I - dd^T) o [3], Hk = sum d_raw r^T [9]  -> 18 (+count)
  double acc[18];
  for (int i = 0; i < 18; ++i) acc[i] = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dx = base[i], dy = base[plane + i], dz = base[2 * plane + i];
    double mx = base[3 * plane + i], my = base[4 * plane + i], mz = base[5 * plane + i];
    const double nd = sqrt(dx * dx + dy * dy + dz * dz);
    mx /= nd; my /= nd; mz /= nd;
    // origin = direction x (moment/|d|)   (geometry_utils.py:322-328)
    const double ox = dy * mz - dz * my, oy = dz * mx - dx * mz, oz = dx * my - dy * mx;
    const double ndc = nd > 1e-12 ? nd : 1e-12;  // F.normalize eps
    const double ux = dx / ndc, uy = dy / ndc, uz = dz / ndc;
    const double p00 = 1 - ux * ux, p01 = -ux * uy, p02 = -ux * uz, p11 = 1 - uy * uy, p12 = -uy * uz, p22 = 1 - uz * uz;
    acc[0] += p00; acc[1] += p01; acc[2] += p02; acc[3] += p11; acc[4] += p12; acc[5] += p22;
    acc[6] += p00 * ox + p01 * oy + p02 * oz;
    acc[7] += p01 * ox + p11 * oy + p12 * oz;
    acc[8] += p02 * ox + p12 * oy + p22 * oz;
    // ideal ray: normalise(K^-1 (col,row,1))
    const double px = (double)(i % w), py = (double)(i / w);
    double rx = Ki[0] * px + Ki[1] * py + Ki[2], ry = Ki[3] * px + Ki[4] * py + Ki[5], rz = Ki[6] * px + Ki[7] * py + Ki[8];
    const double rn = sqrt(rx * rx + ry * ry + rz * rz);
    rx /= rn; ry /= rn; rz /= rn;
    // H = B^T A with B = raw directions, A = ideal rays  (geometry_utils.py:299)
    acc[9] += dx * rx;  acc[10] += dx * ry; acc[11] += dx * rz;
    acc[12] += dy * rx; acc[13] += dy * ry; acc[14] += dy * rz;
    acc[15] += dz * rx; acc[16] += dz * ry; acc[17] += dz * rz;
  }
  block_sum<18>(acc, red);
  if (threadIdx.x == 0) {
    const double M[9] = {acc[0], acc[1], acc[2], acc[1], acc[3], acc[4], acc[2], acc[4], acc[5]};
    double Mi[9];
    inv3(M, Mi);
    double c[3];
    for (int i = 0; i < 3; ++i) c[i] = Mi[i * 3] * acc[6] + Mi[i * 3 + 1] * acc[7] + Mi[i * 3 + 2] * acc[8];
    double U[9], S[3], V[9];
    svd3(acc + 9, U, S, V);
    // R = U diag(1,1,sign det(U V^T)) V^T ; the function returns R^T = extrinsic rotation
    double Vt[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Vt[i * 3 + j] = V[j * 3 + i];
    double UVt[9];
    mul3(U, Vt, UVt);
    const double sg = det3(UVt) < 0 ? -1.0 : 1.0;
    double Us[9];
    for (int i = 0; i < 3; ++i) { Us[i * 3] = U[i * 3]; Us[i * 3 + 1] = U[i * 3 + 1]; Us[i * 3 + 2] = U[i * 3 + 2] * sg; }
    double R[9];
    mul3(Us, Vt, R);
    double Re[9];  // R^T
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Re[i * 3 + j] = R[j * 3 + i];
    double tr[3];
    for (int i = 0; i < 3; ++i) tr[i] = -(Re[i * 3] * c[0] + Re[i * 3 + 1] * c[1] + Re[i * 3 + 2] * c[2]);
    float* e = ext + (long long)b * 16 * T;
    float* pz = pose + (long long)b * 16 * T;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double ev = 0.0, pv = 0.0;
        if (i < 3 && j < 3) { ev = Re[i * 3 + j]; pv = Re[j * 3 + i]; }
        else if (i < 3 && j == 3) { ev = tr[i]; pv = c[i]; }  // inverse of [Re | -Re c] is [Re^T | c]
        else if (i == 3 && j == 3) { ev = 1.0; pv = 1.0; }
        e[(i * 4 + j) * T + t] = (float)ev;
        pz[(i * 4 + j) * T + t] = (float)pv;
      }
    for (int i = 0; i < 3; ++i) centers[((long long)b * T + t) * 3 + i] = (float)c[i];
  }
}

// ---------------------------------------------------------------------------------------------- affine aligner
// moments over n elements per batch row: sum x, sum y, sum xx, sum xy, count with x = f(pred), y = f(target)
__global__ void affine_moments_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n,
                                      long long pred_stride, long long target_stride, int inverse,
                                      double* __restrict__ mom /* [B,5] zero-initialised */) {
  __shared__ double red[5 * 32];
  const int b = blockIdx.y;
  const float* p = pred + b * pred_stride;
  const float* q = target + b * target_stride;
  double a[5] = {0, 0, 0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float x = p[i], y = q[i];
    if (inverse) {
      x = x > 0.f ? 1.0f / x : 0.f;
      y = y > 0.f ? 1.0f / y : 0.f;
    }
    a[0] += x; a[1] += y; a[2] += (double)x * x; a[3] += (double)x * y; a[4] += 1.0;
  }
  block_sum<5>(a, red);
  if (threadIdx.x == 0)
    for (int i = 0; i < 5; ++i) atomicAdd(&mom[b * 5 + i], a[i]);
}
__global__ void affine_solve_kernel(const double* __restrict__ mom, float* __restrict__ sol, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double sx = mom[b * 5], sy = mom[b * 5 + 1], sxx = mom[b * 5 + 2], sxy = mom[b * 5 + 3], n = mom[b * 5 + 4];
  const double det = sxx * n - sx * sx;
  double s = 1.0, t = 0.0;
  if (fabs(det) > 1e-300) {
    s = (sxy * n - sx * sy) / det;
    t = (sxx * sy - sx * sxy) / det;
  }
  sol[b * 2] = (float)s;
  sol[b * 2 + 1] = (float)t;
}
__global__ void affine_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ sol,
                                    long long n, int inverse) {
  const int b = blockIdx.y;
  const float s = sol[b * 2], t = sol[b * 2 + 1];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[b * n + i];
    if (inverse) v = v > 0.f ? 1.0f / v : 0.f;
    v = s * v + t;
    if (inverse) v = v > 0.f ? 1.0f / v : 0.f;
    y[b * n + i] = v;
  }
}

// ---------------------------------------------------------------------------------------------- sim(3) window aligner
// K17: KabaschUmeyama3DAligner.solve (aligner.py:177-237) on the device. Point maps X = pose [depth K^-1 (u,v,1); 1]
// (geometry_utils.py:13-53) of every `step`-th overlap frame are formed on the fly for the current window (src) and
// the stitched buffer (dst); a weighted Umeyama fit is iterated with consensus re-selection
// (|dst - T src| < thr, thr = 0.01 * q98(depth), aligner.py:187-188) instead of the reference's randomised
// skimage RANSAC on a 10% subsample: deterministic, uses every point, and equals the closed-form Umeyama
// solution whenever all points are inliers. The acceptance radius is graduated (2^(iters-1-it) * thr).
struct Sim3Frame {
  double Kinv[9];
  double P[12];  // pose rows 0..2 (3x4)
};
__device__ void sim3_point(const Sim3Frame& f, float depth, int u, int v, double* X) {
  const double cx = (f.Kinv[0] * u + f.Kinv[1] * v + f.Kinv[2]) * depth;
  const double cy = (f.Kinv[3] * u + f.Kinv[4] * v + f.Kinv[5]) * depth;
  const double cz = (f.Kinv[6] * u + f.Kinv[7] * v + f.Kinv[8]) * depth;
  for (int i = 0; i < 3; ++i) X[i] = f.P[i * 4] * cx + f.P[i * 4 + 1] * cy + f.P[i * 4 + 2] * cz + f.P[i * 4 + 3];
}
__device__ void sim3_load_frame(const float* K_b44t, const float* pose_b16t, int T, int t, Sim3Frame& f) {
  double K[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) K[i * 3 + j] = (double)K_b44t[(i * 4 + j) * T + t];
  inv3(K, f.Kinv);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) f.P[i * 4 + j] = (double)pose_b16t[(i * 4 + j) * T + t];
}

// sums[17]: n, sum src(3), sum dst(3), sum |src|^2, sum dst src^T (9). Tprev (16 doubles, row-major) or use_prev = 0.
__global__ void __launch_bounds__(256)
sim3_reduce_kernel(const float* __restrict__ depth_s, const float* __restrict__ K_s, const float* __restrict__ pose_s, int Ts,
                   const float* __restrict__ depth_d, const float* __restrict__ K_d, const float* __restrict__ pose_d, int Td,
                   int nframes, int step, int H, int W, const double* __restrict__ Tprev, int use_prev, const float* thr_ptr,
                   float thr_scale, double* __restrict__ sums) {
  __shared__ double red[17 * 32];
  __shared__ Sim3Frame fs, fd;
  const int fi = blockIdx.y;
  const int t = fi * step;
  if (threadIdx.x == 0) {
    sim3_load_frame(K_s, pose_s, Ts, t, fs);
    sim3_load_frame(K_d, pose_d, Td, t, fd);
  }
  __syncthreads();
  const double thr2 = (double)thr_ptr[0] * thr_ptr[0] * (double)thr_scale * (double)thr_scale;
  double a[17];
  for (int i = 0; i < 17; ++i) a[i] = 0.0;
  const int n = H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int v = p / W, u = p - v * W;
    double S[3], D[3];
    sim3_point(fs, depth_s[(long long)t * n + p], u, v, S);
    sim3_point(fd, depth_d[(long long)t * n + p], u, v, D);
    bool use = true;
    if (use_prev) {
      double e2 = 0.0;
      for (int i = 0; i < 3; ++i) {
        const double m = Tprev[i * 4] * S[0] + Tprev[i * 4 + 1] * S[1] + Tprev[i * 4 + 2] * S[2] + Tprev[i * 4 + 3];
        e2 += (D[i] - m) * (D[i] - m);
      }
      use = e2 < thr2;
    }
    if (use) {
      a[0] += 1.0;
      for (int i = 0; i < 3; ++i) { a[1 + i] += S[i]; a[4 + i] += D[i]; }
      a[7] += S[0] * S[0] + S[1] * S[1] + S[2] * S[2];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) a[8 + i * 3 + j] += D[i] * S[j];
    }
  }
  block_sum<17>(a, red);
  if (threadIdx.x == 0)
    for (int i = 0; i < 17; ++i) atomicAdd(&sums[i], a[i]);
}

// Umeyama (with scale) from the accumulated moments -> T (4x4 row-major doubles) and scale; keeps Tprev if degenerate.
__global__ void sim3_solve_kernel(const double* __restrict__ sums, double* __restrict__ Tout, int min_points) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double n = sums[0];
  if (n < (double)min_points) return;  // keep the previous estimate (initialised to identity by the host)
  double ms[3], md[3];
  for (int i = 0; i < 3; ++i) { ms[i] = sums[1 + i] / n; md[i] = sums[4 + i] / n; }
  double cov[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[i * 3 + j] = sums[8 + i * 3 + j] / n - md[i] * ms[j];
  const double var_s = sums[7] / n - (ms[0] * ms[0] + ms[1] * ms[1] + ms[2] * ms[2]);
  double U[9], S[3], V[9];
  svd3(cov, U, S, V);
  double Vt[9], UVt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Vt[i * 3 + j] = V[j * 3 + i];
  mul3(U, Vt, UVt);
  const double dsg = det3(cov) < 0 ? -1.0 : 1.0;  // Umeyama: reflect when det(cov) < 0
  (void)UVt;
  double Ud[9];
  for (int i = 0; i < 3; ++i) { Ud[i * 3] = U[i * 3]; Ud[i * 3 + 1] = U[i * 3 + 1]; Ud[i * 3 + 2] = U[i * 3 + 2] * dsg; }
  double R[9];
  mul3(Ud, Vt, R);
  const double scale = var_s > 0 ? (S[0] + S[1] + dsg * S[2]) / var_s : 1.0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Tout[i * 4 + j] = scale * R[i * 3 + j];
    Tout[i * 4 + 3] = md[i] - scale * (R[i * 3] * ms[0] + R[i * 3 + 1] * ms[1] + R[i * 3 + 2] * ms[2]);
  }
  Tout[12] = Tout[13] = Tout[14] = 0.0;
  Tout[15] = 1.0;
  Tout[16] = scale;
}

}  // namespace l4p

using namespace l4p;

extern "C" int l4p_pose_from_rays(const float* rays, const float* k_norm, int mode, int B, int T, int h, int w, int outH,
                                  int outW, float reproj_threshold, int refits, double* ws_kgrid, float* ext, float* pose,
                                  float* centers, float* k_est, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  L4P_REQUIRE(rays && ext && pose && centers, L4P_ERR_ARG, "l4p_pose_from_rays: null pointer");
  L4P_REQUIRE(B > 0 && T > 0 && h > 0 && w > 0, L4P_ERR_SHAPE, "l4p_pose_from_rays: bad shape");
  L4P_REQUIRE(mode == 0 || mode == 1, L4P_ERR_ARG, "l4p_pose_from_rays: mode=%d", mode);
  if (mode == 0) {
    L4P_REQUIRE(k_norm != nullptr, L4P_ERR_ARG, "l4p_pose_from_rays: mode 0 needs intrinsics");
  } else {
    L4P_REQUIRE(ws_kgrid && k_est, L4P_ERR_ARG, "l4p_pose_from_rays: mode 1 needs ws_kgrid and k_est");
    L4P_REQUIRE(h * w <= 48 * 1024, L4P_ERR_SHAPE, "l4p_pose_from_rays: ray grid too large");
    estimate_k_kernel<<<B, 256, (size_t)h * w, stream>>>(rays, T, h, w, outH, outW, reproj_threshold, refits, ws_kgrid,
                                                          k_est);
    L4P_CHECK_CUDA(cudaGetLastError());
  }
  pose_kernel<<<B * T, 256, 0, stream>>>(rays, k_norm, ws_kgrid, mode, T, h, w, ext, pose, centers);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_affine_align_solve(const float* pred, const float* target, int B, int64_t n, int64_t pred_stride,
                                      int64_t target_stride, int inverse, double* ws_moments, float* sol, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  L4P_REQUIRE(pred && target && ws_moments && sol, L4P_ERR_ARG, "l4p_affine_align_solve: null pointer");
  L4P_REQUIRE(B > 0 && n > 0, L4P_ERR_SHAPE, "l4p_affine_align_solve: empty input");
  L4P_CHECK_CUDA(cudaMemsetAsync(ws_moments, 0, sizeof(double) * 5 * B, stream));
  long long g = (n + 255) / 256;
  if (g > 4LL * host_num_sms()) g = 4LL * host_num_sms();
  affine_moments_kernel<<<dim3((unsigned)g, B), 256, 0, stream>>>(pred, target, n, pred_stride, target_stride, inverse,
                                                                  ws_moments);
  L4P_CHECK_CUDA(cudaGetLastError());
  affine_solve_kernel<<<(B + 63) / 64, 64, 0, stream>>>(ws_moments, sol, B);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_affine_align_apply(const float* x, float* y, const float* sol, int B, int64_t n, int inverse,
                                      void* stream_) {
  L4P_REQUIRE(x && y && sol, L4P_ERR_ARG, "l4p_affine_align_apply: null pointer");
  L4P_REQUIRE(B > 0 && n > 0, L4P_ERR_SHAPE, "l4p_affine_align_apply: empty input");
  long long g = (n + 255) / 256;
  if (g > 8LL * host_num_sms()) g = 8LL * host_num_sms();
  affine_apply_kernel<<<dim3((unsigned)g, B), 256, 0, (cudaStream_t)stream_>>>(x, y, sol, n, inverse);
  L4P_CHECK_CUDA(cudaGetLastError());
  return L4P_OK;
}

extern "C" int l4p_sim3_align(const float* depth_src, const float* K_src, const float* pose_src, int T_src,
                              const float* depth_dst, const float* K_dst, const float* pose_dst, int T_dst, int overlap,
                              int frame_step, int H, int W, const float* thr_dev, int iters, int min_points,
                              double* ws /* 17 + 17 doubles */, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  L4P_REQUIRE(depth_src && K_src && pose_src && depth_dst && K_dst && pose_dst && thr_dev && ws, L4P_ERR_ARG,
              "l4p_sim3_align: null pointer");
  L4P_REQUIRE(overlap > 0 && frame_step > 0 && H > 0 && W > 0 && iters >= 1 && overlap <= T_src && overlap <= T_dst,
              L4P_ERR_SHAPE, "l4p_sim3_align: bad shape");
  const int nframes = (overlap + frame_step - 1) / frame_step;
  double* sums = ws;
  double* Tcur = ws + 17;  // 16 matrix entries + scale
  const double ident[17] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1};
  L4P_CHECK_CUDA(cudaMemcpyAsync(Tcur, ident, sizeof(ident), cudaMemcpyHostToDevice, stream));
  int gx = (H * W + 255) / 256;
  if (gx > 64) gx = 64;
  for (int it = 0; it < iters; ++it) {
    // graduated consensus: the acceptance radius shrinks by 2x per iteration down to the reference threshold, so a
    // biased all-points start still keeps the true inliers while gross outliers drop out first
    const int sh = iters - 1 - it;
    const float thr_scale = (float)(1u << (sh < 0 ? 0 : (sh > 20 ? 20 : sh)));
    L4P_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 17, stream));
    sim3_reduce_kernel<<<dim3(gx, nframes), 256, 0, stream>>>(depth_src, K_src, pose_src, T_src, depth_dst, K_dst, pose_dst,
                                                              T_dst, nframes, frame_step, H, W, Tcur, it > 0, thr_dev,
                                                              thr_scale, sums);
    L4P_CHECK_CUDA(cudaGetLastError());
    sim3_solve_kernel<<<1, 32, 0, stream>>>(sums, Tcur, min_points);
    L4P_CHECK_CUDA(cudaGetLastError());
  }
  return L4P_OK;
}
