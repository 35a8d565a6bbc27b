// Training-step kernels: batch normalisation, RaGAN / content / topographic / SSIM losses with
// their gradients, PSNR partial sums and the Chainer-variant Adam update.
// Reference semantics: srgan_train.py:636-689 (BN + LeakyReLU), :841-1009 (losses),
// :1043-1048 (Adam); third-party details restated in SURVEY App. B.7, B.9, B.10.
#include "common.cuh"

namespace dbm {

__device__ __forceinline__ float block_sum(float s, float* red) {
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}

// ---- BatchNormalization(axis=(0,2,3), eps=1e-5, decay=0.9) -------------------------------------
// train: batch mean / biased variance, running stats updated with the unbiased variance; eval: running stats.
// Writes mean[c], invstd[c] for apply/backward.
// Per-channel reductions are split over the batch (grid = C x S blocks, S ~ 4 blocks per SM / C): one block per
// channel left 64-channel layers on 64 of 148 SMs. Partials are double and are combined in a fixed order by the
// last block of the channel to arrive (deterministic); scratch is a library-owned static buffer with one slot per
// stream that has ever issued a BN call (kBnSlots of them), so BN calls on different streams -- a second
// discriminator, a D pass overlapping another -- never share partials or counters; calls on ONE stream are ordered.
constexpr int kBnMaxC = 1024, kBnMaxSplit = 32, kBnSlots = 8;
__device__ double g_bn_part_all[kBnSlots][kBnMaxC * kBnMaxSplit * 2];
__device__ unsigned int g_bn_count_all[kBnSlots][kBnMaxC];
// grouped calls (G independent BatchNormalization passes over one stacked batch in ONE launch): per (group, channel)
// batch variance and a per-channel counter of finished groups -- the group that finishes last applies the running-
// statistics updates / gradient accumulations of all groups in group order, exactly as G consecutive calls would
__device__ float g_bn_var_all[kBnSlots][kBnMaxC];
__device__ unsigned int g_bn_gcount_all[kBnSlots][kBnMaxC];
// true in the last group of channel c to get here (everything the other groups wrote before is visible: read with __ldcg)
__device__ __forceinline__ bool bn_last_group(unsigned int* gcount, int c, int G) {
  __threadfence();
  const bool last = atomicAdd(&gcount[c], 1u) == (unsigned)(G - 1);
  if (last) {
    gcount[c] = 0;
    __threadfence();
  }
  return last;
}

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];   // fixed order, every thread gets the total
  return t;
}
// returns true in the LAST block of channel c to finish (its view of every partial is complete)
__device__ __forceinline__ bool bn_publish(double* g_bn_part, unsigned int* g_bn_count, int c, int split, int S,
                                           double a, double b) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    g_bn_part[((size_t)c * kBnMaxSplit + split) * 2] = a;
    g_bn_part[((size_t)c * kBnMaxSplit + split) * 2 + 1] = b;
    __threadfence();
    last = atomicAdd(&g_bn_count[c], 1u) == (unsigned)(S - 1);
    if (last) {
      g_bn_count[c] = 0;   // ready for the next call
      __threadfence();
    }
  }
  __syncthreads();
  return last;
}

// scratch slot of a stream (assigned on first use; -1 when more than kBnSlots streams issue BN calls)
static int bn_slot(cudaStream_t st) {
  static cudaStream_t owners[kBnSlots];
  static int used = 0;
  for (int i = 0; i < used; ++i)
    if (owners[i] == st) return i;
  if (used == kBnSlots) return -1;
  owners[used] = st;
  return used++;
}

static int bn_split(int n, int c) {
  int s = (4 * num_sms() + c - 1) / c;
  if (s > n) s = n;
  if (s > kBnMaxSplit) s = kBnMaxSplit;
  return s < 1 ? 1 : s;
}

__global__ void bn_stats_kernel(const float* __restrict__ x, int N, int C, int HW, float eps, float decay, int train,
                                float* __restrict__ avg_mean, float* __restrict__ avg_var,
                                float* __restrict__ mean_out, float* __restrict__ invstd_out, int slot) {
  // N = samples PER GROUP; group g = blockIdx.z owns samples [g N, (g + 1) N) and rows g of mean_out / invstd_out
  __shared__ double red[32];
  double* g_bn_part = g_bn_part_all[slot];
  unsigned int* g_bn_count = g_bn_count_all[slot];
  const int c = blockIdx.x, split = blockIdx.y, S = gridDim.y, g = blockIdx.z, G = gridDim.z;
  const int cg = g * C + c;
  if (!train) {
    if (threadIdx.x == 0 && split == 0) {
      mean_out[cg] = avg_mean[c];
      invstd_out[cg] = rsqrtf(avg_var[c] + eps);
    }
    return;
  }
  x += (size_t)g * N * C * HW;
  const long n0 = (long)N * split / S, n1 = (long)N * (split + 1) / S;
  const long cnt = (n1 - n0) * HW;
  double s = 0.0, q = 0.0;   // sum and sum of squares in double: var = E[x^2] - mean^2 without cancellation trouble
  for (int i = threadIdx.x; i < (int)cnt; i += blockDim.x) {
    const int n = i / HW, r = i - n * HW;
    const double v = (double)x[((n0 + n) * C + c) * HW + r];
    s += v;
    q += v * v;
  }
  s = block_sum_d(s, red);
  q = block_sum_d(q, red);
  if (!bn_publish(g_bn_part, g_bn_count, cg, split, S, s, q)) return;
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < S; ++k) {
      ts += g_bn_part[((size_t)cg * kBnMaxSplit + k) * 2];
      tq += g_bn_part[((size_t)cg * kBnMaxSplit + k) * 2 + 1];
    }
    const double m = (double)N * HW;
    const double mean = ts / m;
    double var = tq / m - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_out[cg] = (float)mean;
    invstd_out[cg] = (float)(1.0 / sqrt(var + (double)eps));
    const float adjust = (float)(m / fmax(m - 1.0, 1.0));
    // explicit roundings: the single-group and the grouped path must produce the same bits (no compiler-chosen fma)
    auto running = [decay](float avg, float v) { return __fmaf_rn(decay, avg, __fmul_rn(1.f - decay, v)); };
    if (G == 1) {
      avg_mean[c] = running(avg_mean[c], (float)mean);
      avg_var[c] = running(avg_var[c], __fmul_rn((float)var, adjust));
    } else {
      g_bn_var_all[slot][cg] = (float)var;
      if (bn_last_group(g_bn_gcount_all[slot], c, G)) {
        float am = avg_mean[c], av = avg_var[c];
        for (int gg = 0; gg < G; ++gg) {   // group order = the order of G consecutive calls
          am = running(am, __ldcg(mean_out + gg * C + c));
          av = running(av, __fmul_rn(__ldcg(&g_bn_var_all[slot][gg * C + c]), adjust));
        }
        avg_mean[c] = am;
        avg_var[c] = av;
      }
    }
  }
}
// y = lrelu(gamma * (x - mean) * invstd + beta)
__global__ void bn_apply_lrelu_kernel(const float* __restrict__ x, float* __restrict__ y,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ mean, const float* __restrict__ invstd, int C, int HW,
                                      long total, long group_elems) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (i / HW) % C;
    const int cg = (int)(i / group_elems) * C + c;
    y[i] = lrelu(gamma[c] * (x[i] - mean[cg]) * invstd[cg] + beta[c]);
  }
}
// Backward of y = lrelu(BN_train(x)): per-channel reductions (dgamma += , dbeta +=) ...
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ dy, int N, int C, int HW,
                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     float* __restrict__ scratch, int slot) {
  // N = samples PER GROUP (group = blockIdx.z); scratch [G][2 C]: sum_dz | sum_dz_xhat of every group
  __shared__ double red[32];
  double* g_bn_part = g_bn_part_all[slot];
  unsigned int* g_bn_count = g_bn_count_all[slot];
  const int c = blockIdx.x, split = blockIdx.y, S = gridDim.y, g = blockIdx.z, G = gridDim.z;
  const int cg = g * C + c;
  const size_t goff = (size_t)g * N * C * HW;
  x += goff; y += goff; dy += goff;
  const long n0 = (long)N * split / S, n1 = (long)N * (split + 1) / S;
  const long cnt = (n1 - n0) * HW;
  const float mu = mean[cg], is = invstd[cg];
  double s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < (int)cnt; i += blockDim.x) {
    const int n = i / HW, r = i - n * HW;
    const long idx = ((n0 + n) * C + c) * HW + r;
    float dz = dy[idx];
    if (y[idx] < 0.f) dz *= kLreluSlope;
    s1 += (double)dz;
    s2 += (double)(dz * (x[idx] - mu) * is);
  }
  s1 = block_sum_d(s1, red);
  s2 = block_sum_d(s2, red);
  if (!bn_publish(g_bn_part, g_bn_count, cg, split, S, s1, s2)) return;
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < S; ++k) {
      t1 += g_bn_part[((size_t)cg * kBnMaxSplit + k) * 2];
      t2 += g_bn_part[((size_t)cg * kBnMaxSplit + k) * 2 + 1];
    }
    float* sum_dz = scratch + (size_t)g * 2 * C;
    sum_dz[c] = (float)t1;
    sum_dz[C + c] = (float)t2;
    if (G == 1) {
      dbeta[c] += (float)t1;
      dgamma[c] += (float)t2;
    } else if (bn_last_group(g_bn_gcount_all[slot], c, G)) {
      float db = dbeta[c], dg = dgamma[c];
      for (int gg = 0; gg < G; ++gg) {   // group order = the order of G consecutive calls
        db = __fadd_rn(db, __ldcg(scratch + (size_t)gg * 2 * C + c));
        dg = __fadd_rn(dg, __ldcg(scratch + (size_t)gg * 2 * C + C + c));
      }
      dbeta[c] = db;
      dgamma[c] = dg;
    }
  }
}
// ... then dx = gamma * invstd * (dz - sum_dz/m - xhat * sum_dz_xhat/m)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ dy, float* __restrict__ dx,
                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                    const float* __restrict__ invstd, const float* __restrict__ scratch, int C, int HW,
                                    long total, long group_elems, float inv_m) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (i / HW) % C;
    const int g = (int)(i / group_elems);
    const int cg = g * C + c;
    const float* sum_dz = scratch + (size_t)g * 2 * C;
    float dz = dy[i];
    if (y[i] < 0.f) dz *= kLreluSlope;
    const float xhat = (x[i] - mean[cg]) * invstd[cg];
    dx[i] = gamma[c] * invstd[cg] * (dz - sum_dz[c] * inv_m - xhat * sum_dz[C + c] * inv_m);
  }
}

// ---- RaGAN sigmoid cross-entropy (srgan_train.py:960-1009) ---------------------------------------
// out[0] = loss, out[1] = binary accuracy of [real; fake] vs [1; 0] (srgan_train.py:1156-1158).
// Optional gradients d_real, d_fake (N each), scaled by `gscale`.
__device__ __forceinline__ float sce(float x, float t) {
  return -(x * (t - (x >= 0.f ? 1.f : 0.f)) - log1pf(expf(-fabsf(x))));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void ragan_loss_kernel(const float* __restrict__ real, const float* __restrict__ fake, int n, float t_rmf,
                                  float t_fmr, float gscale, float* __restrict__ out, float* __restrict__ d_real,
                                  float* __restrict__ d_fake) {
  __shared__ float red[32];
  float sr = 0.f, sf = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sr += real[i];
    sf += fake[i];
  }
  const float rbar = block_sum(sr, red) / n;
  const float fbar = block_sum(sf, red) / n;
  float l = 0.f, acc = 0.f, g1 = 0.f, g2 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = real[i] - fbar, b = fake[i] - rbar;
    l += sce(a, t_rmf) + sce(b, t_fmr);
    acc += (real[i] >= 0.f ? 1.f : 0.f) + (fake[i] >= 0.f ? 0.f : 1.f);
    g1 += sigmoidf(a) - t_rmf;
    g2 += sigmoidf(b) - t_fmr;
  }
  l = block_sum(l, red);
  acc = block_sum(acc, red);
  g1 = block_sum(g1, red);
  g2 = block_sum(g2, red);
  if (threadIdx.x == 0) {
    out[0] = l / n;
    out[1] = acc / (2.f * n);
  }
  if (d_real && d_fake) {
    const float inv = 1.f / n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float a = real[i] - fbar, b = fake[i] - rbar;
      d_real[i] = gscale * inv * ((sigmoidf(a) - t_rmf) - inv * g2);
      d_fake[i] = gscale * inv * ((sigmoidf(b) - t_fmr) - inv * g1);
    }
  }
}

// ---- generator image losses: content L1, topographic L1 (4x4 avg-pool), SSIM 9x9 valid ----------
// (srgan_train.py:871, 882-887, 932-956). One block per image (1 channel, H x W, H,W <= 48).
// sums[0..3] += sum|yp-yt|, sum|pool(yp)-xt|, sum ssim_map, sum (yp-yt)^2.
// dy (optional) = w_content*dL1 + w_topo*dTopo + w_struct*d(1-ssim), all with mean normalisers.
constexpr int kMaxImg = 48;
constexpr int kWin = 9;
__constant__ float c_gauss[kWin];  // normalised 1-D Gaussian, sigma 1.5

constexpr int kLossMaxN = 65536;
__device__ float g_loss_part[kLossMaxN * 4];   // per-image partial sums of gen_image_loss_kernel
__device__ unsigned int g_loss_count;

__global__ void gen_image_loss_kernel(const float* __restrict__ yp, const float* __restrict__ yt,
                                      const float* __restrict__ xt, int N, int H, int W, float w_content,
                                      float w_topo, float w_struct, float* __restrict__ sums,
                                      float* __restrict__ dy) {
  __shared__ float sp[kMaxImg * kMaxImg], st[kMaxImg * kMaxImg];
  __shared__ float gP[(kMaxImg - 8) * (kMaxImg - 8)], gQ[(kMaxImg - 8) * (kMaxImg - 8)],
      gR[(kMaxImg - 8) * (kMaxImg - 8)];
  __shared__ float red[32];
  const int n = blockIdx.x;
  const int HW = H * W;
  const int MH = H - kWin + 1, MW = W - kWin + 1;
  const int PH = H / 4, PW = W / 4;
  const float* p = yp + (long)n * HW;
  const float* q = yt + (long)n * HW;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    sp[i] = p[i];
    st[i] = q[i];
  }
  __syncthreads();
  const float C1 = 1e-4f, C2 = 9e-4f;
  const float ssim_scale = -w_struct / ((float)N * MH * MW);  // d(1 - mean ssim)
  float s_ssim = 0.f;
  for (int wi = threadIdx.x; wi < MH * MW; wi += blockDim.x) {
    const int wy = wi / MW, wx = wi - wy * MW;
    float P = 0.f, M2 = 0.f, Q = 0.f, S2 = 0.f, R = 0.f;
    for (int a = 0; a < kWin; ++a) {
      float rp = 0.f, rt = 0.f, rq = 0.f, rs = 0.f, rr = 0.f;
      for (int b = 0; b < kWin; ++b) {
        const float g = c_gauss[b];
        const float u = sp[(wy + a) * W + wx + b], v = st[(wy + a) * W + wx + b];
        rp += g * u; rt += g * v; rq += g * u * u; rs += g * v * v; rr += g * u * v;
      }
      const float g = c_gauss[a];
      P += g * rp; M2 += g * rt; Q += g * rq; S2 += g * rs; R += g * rr;
    }
    const float s1 = Q - P * P, s2 = S2 - M2 * M2, s12 = R - P * M2;
    const float A = 2.f * P * M2 + C1, B = 2.f * s12 + C2, Cc = P * P + M2 * M2 + C1, D = s1 + s2 + C2;
    const float inv = 1.f / (Cc * D);
    s_ssim += A * B * inv;
    // partial derivatives wrt P = G*yp, Q = G*yp^2, R = G*(yp*yt)
    const float dP = (2.f * M2 * B - 2.f * M2 * A) * inv - A * B * inv * inv * (2.f * P * D - 2.f * P * Cc);
    const float dQ = -A * B * inv / D;
    const float dR = 2.f * A * inv;
    gP[wi] = ssim_scale * dP;
    gQ[wi] = ssim_scale * dQ;
    gR[wi] = ssim_scale * dR;
  }
  __syncthreads();
  float s_l1 = 0.f, s_sq = 0.f;
  const float l1_scale = w_content / ((float)N * HW);
  const float topo_scale = w_topo / ((float)N * PH * PW) / 16.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float u = sp[i], v = st[i];
    const float d = u - v;
    s_l1 += fabsf(d);
    s_sq += d * d;
    if (dy) {
      float g = l1_scale * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      // topographic term: pooled cell (y/4, x/4)
      if (y / 4 < PH && x / 4 < PW) {
        float pool = 0.f;
        const int by = (y / 4) * 4, bx = (x / 4) * 4;
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b) pool += sp[(by + a) * W + bx + b];
        const float dd = pool * (1.f / 16.f) - xt[((long)n * PH + y / 4) * PW + x / 4];
        g += topo_scale * (dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f));
      }
      // SSIM term: adjoint of the valid Gaussian filtering
      float gp = 0.f, gq = 0.f, gr = 0.f;
      for (int a = 0; a < kWin; ++a) {
        const int wy = y - a;
        if (wy < 0 || wy >= MH) continue;
        for (int b = 0; b < kWin; ++b) {
          const int wx = x - b;
          if (wx < 0 || wx >= MW) continue;
          const float gg = c_gauss[a] * c_gauss[b];
          gp += gg * gP[wy * MW + wx];
          gq += gg * gQ[wy * MW + wx];
          gr += gg * gR[wy * MW + wx];
        }
      }
      g += gp + 2.f * u * gq + v * gr;
      dy[(long)n * HW + i] = g;
    }
  }
  float s_topo = 0.f;
  for (int i = threadIdx.x; i < PH * PW; i += blockDim.x) {
    const int cy = i / PW, cx = i - cy * PW;
    float pool = 0.f;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) pool += sp[(cy * 4 + a) * W + cx * 4 + b];
    s_topo += fabsf(pool * (1.f / 16.f) - xt[(long)n * PH * PW + i]);
  }
  s_l1 = block_sum(s_l1, red);
  s_topo = block_sum(s_topo, red);
  s_ssim = block_sum(s_ssim, red);
  s_sq = block_sum(s_sq, red);
  // per-image partials, summed in image order by the last block to arrive: the four sums are the same bits every run
  __shared__ bool last;
  if (threadIdx.x == 0) {
    float* part = g_loss_part + (size_t)n * 4;
    part[0] = s_l1; part[1] = s_topo; part[2] = s_ssim; part[3] = s_sq;
    __threadfence();
    last = atomicAdd(&g_loss_count, 1u) == (unsigned)(N - 1);
    if (last) g_loss_count = 0;
  }
  __syncthreads();
  if (last && threadIdx.x < 4) {
    __threadfence();
    double t = 0.0;
    for (int k = 0; k < N; ++k) t += (double)g_loss_part[(size_t)k * 4 + threadIdx.x];
    sums[threadIdx.x] = (float)t;
  }
}

// ---- chainer.optimizers.Adam: eps added to the UNcorrected sqrt(v) -------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr_t, float beta1, float beta2, float eps,
                            float grad_scale) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);
    const float vi = v[i] + (1.f - beta2) * (gi * gi - v[i]);
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

// Device-side step counter variant (CUDA-graph replays: a host-computed bias correction would be frozen into the graph)
__global__ void adam_tick_kernel(int* __restrict__ t) { *t += 1; }
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long n, double alpha, float beta1, float beta2, float eps,
                                const int* __restrict__ t_dev, float grad_scale) {
  const int t = *t_dev;
  const double fix1 = 1.0 - pow((double)beta1, (double)t), fix2 = 1.0 - pow((double)beta2, (double)t);
  const float lr_t = (float)(alpha * sqrt(fix2) / fix1);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = m[i] + (1.f - beta1) * (gi - m[i]);
    const float vi = v[i] + (1.f - beta2) * (gi * gi - v[i]);
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace dbm

using namespace dbm;

static inline int grid_for(long total) {
  long b = (total + 255) / 256;
  long cap = (long)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// `groups` independent BatchNormalization passes over a batch stacked along N (group g = samples [g n, (g + 1) n)) in
// ONE pair of launches: batch statistics per group (save_mean / save_invstd: [groups][c]), running statistics updated
// group after group -- the values of `groups` consecutive single-group calls, bit for bit.
extern "C" int dbm_bn_lrelu_fwd_groups_f32(const float* x, float* y, const float* gamma, const float* beta,
                                           float* avg_mean, float* avg_var, float* save_mean, float* save_invstd,
                                           int groups, int n, int c, int hw, float eps, float decay, int train,
                                           cudaStream_t st) {
  DBM_REQUIRE(groups > 0 && n > 0 && c > 0 && hw > 0, "bn: empty input");
  DBM_REQUIRE(groups * c <= kBnMaxC && (long)n * hw < (1L << 31),
              "bn: %d x %d channels / %d x %d elements exceed the reduction scratch", groups, c, n, hw);
  const int slot = bn_slot(st);
  DBM_REQUIRE(slot >= 0, "bn: more than %d streams issue BatchNormalization calls", kBnSlots);
  bn_stats_kernel<<<dim3(c, train ? bn_split(n, c) : 1, groups), 256, 0, st>>>(
      x, n, c, hw, eps, decay, train, avg_mean, avg_var, save_mean, save_invstd, slot);
  int rc = check_launch("bn_stats");
  if (rc) return rc;
  const long group_elems = (long)n * c * hw, total = group_elems * groups;
  bn_apply_lrelu_kernel<<<grid_for(total), 256, 0, st>>>(x, y, gamma, beta, save_mean, save_invstd, c, hw, total,
                                                         group_elems);
  return check_launch("bn_apply_lrelu");
}

extern "C" int dbm_bn_lrelu_fwd_f32(const float* x, float* y, const float* gamma, const float* beta, float* avg_mean,
                                    float* avg_var, float* save_mean, float* save_invstd, int n, int c, int hw,
                                    float eps, float decay, int train, cudaStream_t st) {
  return dbm_bn_lrelu_fwd_groups_f32(x, y, gamma, beta, avg_mean, avg_var, save_mean, save_invstd, 1, n, c, hw, eps,
                                     decay, train, st);
}

// backward of the grouped call: scratch [groups][2 c]; dgamma / dbeta accumulate the groups' sums in group order
extern "C" int dbm_bn_lrelu_bwd_groups_f32(const float* x, const float* y, const float* dy, float* dx,
                                           const float* gamma, const float* save_mean, const float* save_invstd,
                                           float* dgamma, float* dbeta, float* scratch, int groups, int n, int c,
                                           int hw, cudaStream_t st) {
  DBM_REQUIRE(groups > 0 && n > 0 && c > 0 && hw > 0, "bn_bwd: empty input");
  DBM_REQUIRE(groups * c <= kBnMaxC && (long)n * hw < (1L << 31),
              "bn_bwd: %d x %d channels / %d x %d elements exceed the reduction scratch", groups, c, n, hw);
  const int slot = bn_slot(st);
  DBM_REQUIRE(slot >= 0, "bn_bwd: more than %d streams issue BatchNormalization calls", kBnSlots);
  bn_bwd_reduce_kernel<<<dim3(c, bn_split(n, c), groups), 256, 0, st>>>(x, y, dy, n, c, hw, save_mean,
                                                                                 save_invstd, dgamma, dbeta, scratch, slot);
  int rc = check_launch("bn_bwd_reduce");
  if (rc) return rc;
  const long group_elems = (long)n * c * hw, total = group_elems * groups;
  bn_bwd_apply_kernel<<<grid_for(total), 256, 0, st>>>(x, y, dy, dx, gamma, save_mean, save_invstd, scratch, c, hw,
                                                       total, group_elems, 1.f / ((float)n * hw));
  return check_launch("bn_bwd_apply");
}

extern "C" int dbm_bn_lrelu_bwd_f32(const float* x, const float* y, const float* dy, float* dx, const float* gamma,
                                    const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                                    float* scratch2c, int n, int c, int hw, cudaStream_t st) {
  return dbm_bn_lrelu_bwd_groups_f32(x, y, dy, dx, gamma, save_mean, save_invstd, dgamma, dbeta, scratch2c, 1, n, c, hw,
                                     st);
}

extern "C" int dbm_ragan_loss_f32(const float* real_pred, const float* fake_pred, int n, float t_real_minus_fake,
                                  float t_fake_minus_real, float grad_scale, float* out2, float* d_real,
                                  float* d_fake, cudaStream_t st) {
  DBM_REQUIRE(n > 0, "ragan_loss: empty batch");
  ragan_loss_kernel<<<1, 256, 0, st>>>(real_pred, fake_pred, n, t_real_minus_fake, t_fake_minus_real, grad_scale,
                                       out2, d_real, d_fake);
  return check_launch("ragan_loss");
}

extern "C" int dbm_gen_image_loss_f32(const float* y_pred, const float* y_true, const float* x_topo, int n, int h,
                                      int w, float w_content, float w_topo, float w_struct, float* sums4,
                                      float* dy, cudaStream_t st) {
  DBM_REQUIRE(n > 0, "gen_image_loss: empty batch");
  DBM_REQUIRE(h >= kWin && w >= kWin && h <= kMaxImg && w <= kMaxImg && h % 4 == 0 && w % 4 == 0,
              "gen_image_loss: image %dx%d unsupported (need 9..48, multiple of 4)", h, w);
  static bool init_dev[64] = {};   // __constant__ memory is per device
  int dev = 0;
  DBM_CUDA(cudaGetDevice(&dev));
  bool& init = init_dev[dev & 63];
  if (!init) {
    float g[kWin];
    double s = 0;
    for (int i = 0; i < kWin; ++i) {
      double d = i - kWin / 2;
      g[i] = (float)exp(-(d * d) / (2.0 * 1.5 * 1.5));
      s += g[i];
    }
    for (int i = 0; i < kWin; ++i) g[i] = (float)(g[i] / s);
    DBM_CUDA(cudaMemcpyToSymbol(c_gauss, g, sizeof(g)));
    init = true;
  }
  DBM_REQUIRE(n <= kLossMaxN, "gen_image_loss: batch %d exceeds %d", n, kLossMaxN);
  DBM_CUDA(cudaMemsetAsync(sums4, 0, 4 * sizeof(float), st));
  gen_image_loss_kernel<<<n, 256, 0, st>>>(y_pred, y_true, x_topo, n, h, w, w_content, w_topo, w_struct, sums4, dy);
  return check_launch("gen_image_loss");
}

extern "C" int dbm_adam_step_f32(float* params, const float* grads, float* m, float* v, long n, float alpha,
                                 float beta1, float beta2, float eps, int t, float grad_scale, cudaStream_t st) {
  DBM_REQUIRE(n > 0 && t >= 1, "adam: bad arguments (n=%ld, t=%d)", n, t);
  const double fix1 = 1.0 - pow((double)beta1, t), fix2 = 1.0 - pow((double)beta2, t);
  const float lr_t = (float)(alpha * sqrt(fix2) / fix1);
  adam_kernel<<<grid_for(n), 256, 0, st>>>(params, grads, m, v, n, lr_t, beta1, beta2, eps, grad_scale);
  return check_launch("adam");
}

extern "C" int dbm_adam_step_dev_f32(float* params, const float* grads, float* m, float* v, long n, float alpha,
                                     float beta1, float beta2, float eps, int* t_dev, float grad_scale,
                                     cudaStream_t st) {
  DBM_REQUIRE(n > 0 && t_dev != nullptr, "adam_dev: bad arguments (n=%ld)", n);
  adam_tick_kernel<<<1, 1, 0, st>>>(t_dev);
  adam_dev_kernel<<<grid_for(n), 256, 0, st>>>(params, grads, m, v, n, (double)alpha, beta1, beta2, eps, t_dev, grad_scale);
  return check_launch("adam_dev");
}
