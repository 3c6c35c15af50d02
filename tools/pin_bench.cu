// Diagnostic: what pinning host memory costs on this box -- cudaHostAlloc against mmap + transparent huge pages +
// cudaHostRegister -- and whether the copies run as fast from either.  nvcc -O2 -o pin_bench tools/pin_bench.cu
#include <chrono>
#include <cstdio>
#include <cstring>
#include <sys/mman.h>
#include <vector>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaFree(0);
  const size_t B = (size_t)72 << 20;
  const int N = 8;
  void* d = nullptr;
  cudaMalloc(&d, B);
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int mode = 0; mode < 3; ++mode) {
    std::vector<void*> p(N);
    double t0 = now();
    for (int i = 0; i < N; ++i) {
      if (mode == 0) {
        cudaHostAlloc(&p[i], B, cudaHostAllocDefault);
      } else {
        p[i] = mmap(nullptr, B, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (mode == 2) madvise(p[i], B, MADV_HUGEPAGE);
        for (size_t o = 0; o < B; o += 4096) ((volatile char*)p[i])[o] = 0;
      }
    }
    double t1 = now();
    if (mode)
      for (int i = 0; i < N; ++i) {
        cudaError_t e = cudaHostRegister(p[i], B, cudaHostRegisterDefault);
        if (e != cudaSuccess) printf("register failed: %s\n", cudaGetErrorString(e));
      }
    double t2 = now();
    cudaMemcpyAsync(d, p[0], B, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    double t3 = now();
    for (int i = 0; i < N; ++i) cudaMemcpyAsync(d, p[i], B, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    double t4 = now();
    for (int i = 0; i < N; ++i) cudaMemcpyAsync(p[i], d, B, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    double t5 = now();
    printf("%s: populate %.1f ms, register %.1f ms (%d x %zu MiB = %.2f GB/s pinned), H2D %.1f GB/s, D2H %.1f GB/s\n",
           mode == 0 ? "cudaHostAlloc        " : mode == 1 ? "mmap 4K + register   " : "mmap THP + register  ", (t1 - t0) * 1e3, (t2 - t1) * 1e3, N,
           B >> 20, N * B / (t2 - t0) / 1e9, N * B / (t4 - t3) / 1e9, N * B / (t5 - t4) / 1e9);
    double t6 = now();
    for (int i = 0; i < N; ++i) {
      if (mode == 0) cudaFreeHost(p[i]);
      else { cudaHostUnregister(p[i]); munmap(p[i], B); }
    }
    printf("   release %.1f ms\n", (now() - t6) * 1e3);
  }
  FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  char buf[128] = {0};
  if (f) { fgets(buf, sizeof buf, f); fclose(f); }
  printf("THP: %s", buf);
  return 0;
}
