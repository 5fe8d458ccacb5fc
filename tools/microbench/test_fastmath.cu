#include <cstdio>
#include <cmath>
#include "../../scikit-optimize_b200/csrc/fastmath.cuh"
__global__ void k(const double* x, double* e, double* s, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { e[i] = skb::exp_nonpos(-x[i]); s[i] = skb::sqrt_nonneg(x[i]); }
}
int main() {
    const int n = 1 << 20;
    double *x, *e, *s;
    cudaMallocManaged(&x, n * 8); cudaMallocManaged(&e, n * 8); cudaMallocManaged(&s, n * 8);
    for (int i = 0; i < n; i++) { double u = (double)rand() / RAND_MAX; x[i] = (i % 4 == 0) ? u * 1e-6 : (i % 4 == 1) ? u * 800 : (i % 4 == 2) ? u * 20 : u * u * 3; }
    x[0] = 0.0; x[1] = 1e-300; x[2] = 745.0; x[3] = 709.0; x[4] = 707.9;
    k<<<n / 256, 256>>>(x, e, s, n);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(err)); return 1; }
    double me = 0, ms = 0;
    for (int i = 0; i < n; i++) {
        double re = exp(-x[i]), rs = sqrt(x[i]);
        double de = re > 1e-300 ? fabs(e[i] - re) / re : fabs(e[i] - re);
        double ds = rs > 0 ? fabs(s[i] - rs) / rs : fabs(s[i]);
        if (de > me) me = de; if (ds > ms) ms = ds;
    }
    printf("max rel err exp %.3e sqrt %.3e  (eps=2.2e-16)  e[0..4]=%g %g %g %g %g s[0]=%g s[1]=%g\n", me, ms, e[0], e[1], e[2], e[3], e[4], s[0], s[1]);
    return 0;
}
