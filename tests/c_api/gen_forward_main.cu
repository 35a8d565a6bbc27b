// C-only host of the generator (no Python, no model.py): creates a dbm_gen handle, loads every parameter array through
// dbm_gen_set_param by its Chainer .npz key, runs dbm_gen_forward and writes the prediction.
//   usage: gen_forward_main <num_residual_blocks> <residual_scaling> <n> <h> <w> <params.bin> <inputs.bin> <out.bin>
//          [precision: 0 = bf16 (default), 1 = bf16x3]
//   params.bin : the arrays of dbm_gen_array_info, in its order, float32, concatenated
//   inputs.bin : x (n,1,h,w) | w1 (n,1,10h,10w) | w2 (n,2,2h,2w) | w3 (n,1,h,w), float32
//   out.bin    : y (n,1,4(h-2),4(w-2)) float32
// Built and run by tests/test_gpu_c_api.py.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "deepbedmap_b200.h"

#define CHECK(call)                                                      \
  do {                                                                   \
    int rc_ = (call);                                                    \
    if (rc_ != 0) {                                                      \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, dbm_last_error()); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

static std::vector<float> read_all(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fseek(f, 0, SEEK_END);
  long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<float> v(bytes / 4);
  if (fread(v.data(), 4, v.size(), f) != v.size()) { fprintf(stderr, "short read of %s\n", path); exit(2); }
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc != 9 && argc != 10) { fprintf(stderr, "usage: see the header of this file\n"); return 2; }
  const int precision = argc == 10 ? atoi(argv[9]) : 0;   // 0 = bf16, 1 = bf16x3 (dbm_gen_set_precision)
  const int nb = atoi(argv[1]);
  const float beta = (float)atof(argv[2]);
  const int n = atoi(argv[3]), h = atoi(argv[4]), w = atoi(argv[5]);
  dbm_gen* gen = nullptr;
  CHECK(dbm_gen_create(nb, beta, 32, &gen));
  CHECK(dbm_gen_set_precision(gen, precision));
  std::vector<float> params = read_all(argv[6]);
  if ((long)params.size() != dbm_gen_count_params(gen)) {
    fprintf(stderr, "params.bin holds %zu floats, the model has %ld\n", params.size(), dbm_gen_count_params(gen));
    return 1;
  }
  for (int i = 0; i < dbm_gen_num_arrays(gen); ++i) {
    const char* key;
    int ndim, dims[4];
    long off;
    CHECK(dbm_gen_array_info(gen, i, &key, &ndim, dims, &off));
    CHECK(dbm_gen_set_param(gen, key, params.data() + off, ndim, dims));
  }
  // error convention: a wrong shape or key is refused with DBM_ERR_INVALID and a message
  int bad[1] = {7};
  if (dbm_gen_set_param(gen, "pre_residual_conv_layer/b", params.data(), 1, bad) != DBM_ERR_INVALID ||
      dbm_gen_set_param(gen, "no_such_layer/W", params.data(), 1, bad) != DBM_ERR_INVALID) {
    fprintf(stderr, "bad set_param calls were not refused\n");
    return 1;
  }
  std::vector<float> in = read_all(argv[7]);
  const size_t nx = (size_t)n * h * w, n1 = 100 * nx, n2 = 8 * nx, ny = (size_t)n * 16 * (h - 2) * (w - 2);
  if (in.size() != nx + n1 + n2 + nx) { fprintf(stderr, "inputs.bin has the wrong size\n"); return 1; }
  float *dx, *d1, *d2, *d3, *dy;
  void* ws;
  const size_t wsb = dbm_gen_workspace_bytes(gen, n, h, w);
  cudaMalloc(&dx, nx * 4); cudaMalloc(&d1, n1 * 4); cudaMalloc(&d2, n2 * 4); cudaMalloc(&d3, nx * 4);
  cudaMalloc(&dy, ny * 4);
  if (cudaMalloc(&ws, wsb) != cudaSuccess) { fprintf(stderr, "workspace of %zu bytes failed\n", wsb); return 1; }
  cudaMemcpy(dx, in.data(), nx * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d1, in.data() + nx, n1 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d2, in.data() + nx + n1, n2 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d3, in.data() + nx + n1 + n2, nx * 4, cudaMemcpyHostToDevice);
  cudaStream_t st;
  cudaStreamCreate(&st);
  if (dbm_gen_forward(gen, dx, d1, d2, d3, n, h, w, dy, ws, wsb - 1, st) != DBM_ERR_INVALID) {
    fprintf(stderr, "a short workspace was not refused\n");
    return 1;
  }
  for (int rep = 0; rep < 2; ++rep) CHECK(dbm_gen_forward(gen, dx, d1, d2, d3, n, h, w, dy, ws, wsb, st));
  if (cudaStreamSynchronize(st) != cudaSuccess) { fprintf(stderr, "forward failed on the device\n"); return 1; }
  std::vector<float> y(ny);
  cudaMemcpy(y.data(), dy, ny * 4, cudaMemcpyDeviceToHost);
  FILE* f = fopen(argv[8], "wb");
  fwrite(y.data(), 4, ny, f);
  fclose(f);
  CHECK(dbm_gen_destroy(gen));
  printf("ok: %d arrays, %ld parameters, workspace %zu bytes, %ld kernel launches\n", nb * 150 + 24,
         (long)params.size(), wsb, dbm_launch_count());
  return 0;
}
