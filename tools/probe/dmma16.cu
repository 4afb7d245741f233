// Probe: fragment layout and timing of mma.sync.aligned.m16n8k16 f64 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void probe(const double *A /*16x16*/, const double *B /*16x8 (k x n)*/, double *C /*16x8*/, int variant) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a[8], b[4], c[4] = {0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) {
    int row, col;
    if (variant == 0) { row = g + 8 * (i & 1); col = t + 4 * (i >> 1); }
    else { row = g + 8 * ((i >> 1) & 1); col = t + 4 * (i & 1) + 8 * (i >> 2); }
    a[i] = A[row * 16 + col];
  }
  for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
  dmma16816(c, a, b);
  for (int i = 0; i < 4; ++i) { const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1); C[row * 8 + col] = c[i]; }
}
__global__ void timing(double *out, int iters, int mode, int chains) {
  double a8[8], b4[4], c[8][4];
  for (int i = 0; i < 8; ++i) a8[i] = 1.0 + threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b4[i] = 0.5 + i;
  for (int k = 0; k < 8; ++k) for (int i = 0; i < 4; ++i) c[k][i] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) { for (int k = 0; k < 8; ++k) if (k < chains) dmma16816(c[k], a8, b4); }
    else { for (int k = 0; k < 8; ++k) if (k < chains) dmma884(c[k][0], c[k][1], a8[0], b4[0]); }
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < 8; ++k) for (int i = 0; i < 4; ++i) s += c[k][i];
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0) / iters; out[blockIdx.x * 2 + 1] = s; }
}
int main() {
  double hA[256], hB[128], hC[128], ref[128];
  for (int i = 0; i < 256; ++i) hA[i] = (i * 37 % 101) * 0.25 - 7;
  for (int i = 0; i < 128; ++i) hB[i] = (i * 53 % 89) * 0.5 - 11;
  for (int r = 0; r < 16; ++r) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < 16; ++k) s += hA[r * 16 + k] * hB[k * 8 + n]; ref[r * 8 + n] = s; }
  double *A, *B, *C; cudaMalloc(&A, sizeof hA); cudaMalloc(&B, sizeof hB); cudaMalloc(&C, sizeof hC);
  cudaMemcpy(A, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, sizeof hB, cudaMemcpyHostToDevice);
  for (int v = 0; v < 2; ++v) {
    probe<<<1, 32>>>(A, B, C, v); cudaMemcpy(hC, C, sizeof hC, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(hC[i] - ref[i]));
    printf("variant %d max err %.3e (%s)\n", v, err, cudaGetErrorString(cudaGetLastError()));
  }
  double *out; cudaMalloc(&out, 1024 * 16); double h[4];
  for (int mode = 0; mode < 2; ++mode) for (int chains : {1, 2, 4, 8}) for (int warps : {1, 4, 8, 16}) {
    timing<<<1, 32 * warps>>>(out, 2000, mode, chains); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%s chains/warp %d warps %2d : %.1f cycles per iteration (= %.1f per MMA per warp)\n", mode ? "m8n8k4  " : "m16n8k16", chains, warps, h[0], h[0] / chains);
  }
  return 0;
}
