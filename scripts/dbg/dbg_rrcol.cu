// Stand-alone harness for k3_rr_col (debugging aid, not part of the library): random u, f on a (2^L+1)^3 grid, the
// fused residual+restriction against plain host loops.   nvcc ... -DRRCOL_LXF=66|68 [-DRRCOL_STAGE=1]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../evostencils_b200/csrc/evo_kernels_rrcol.cuh"

using namespace evo;
using namespace evo::star;

static Geom geom(int level)
{
    Geom g;
    g.n = (1 << level) + 1; g.dim = 3; g.pitch = (g.n + 15) / 16 * 16; g.nz = g.n;
    g.plane = (long long)g.pitch * g.n; g.total = g.plane * g.nz;
    g.zlo = g.zin0 = 1; g.zhi = g.zin1 = g.n - 2; g.zpar = 0; g.zoff = 0;
    return g;
}

#ifndef NW_
#define NW_ 6
#define RC_ 2
#define NPS_ 4
#define MINB_ 2
#endif

int main(int argc, char **argv)
{
    const int level = argc > 1 ? atoi(argv[1]) : 6;
    Geom gf = geom(level), gc = geom(level - 1);
    std::vector<double> u(gf.total, 0.0), f(gf.total, 0.0), ref(gc.total, 0.0), out(gc.total, -7.0);
    srand(1);
    for (int z = 0; z < gf.n; ++z) for (int y = 0; y < gf.n; ++y) for (int x = 0; x < gf.n; ++x) {
        const long long i = z * gf.plane + (long long)y * gf.pitch + x;
        u[i] = rand() / (double)RAND_MAX - 0.5;
        f[i] = rand() / (double)RAND_MAX - 0.5;
    }
    Star7 c{-1.0, -1.5, -2.0, 9.5, -2.25, -1.25, -0.75};
    DenseW W;
    for (int q = 0; q < 27; ++q) W.w[q] = 0.01 * (q + 1);
    // host reference
    std::vector<double> r(gf.total, 0.0);
    for (int z = 1; z < gf.n - 1; ++z) for (int y = 1; y < gf.n - 1; ++y) for (int x = 1; x < gf.n - 1; ++x) {
        const long long i = z * gf.plane + (long long)y * gf.pitch + x;
        double s = 0.0;
        s = s + c.zm * u[i - gf.plane]; s = s + c.ym * u[i - gf.pitch]; s = s + c.xm * u[i - 1]; s = s + c.c * u[i];
        s = s + c.xp * u[i + 1]; s = s + c.yp * u[i + gf.pitch]; s = s + c.zp * u[i + gf.plane];
        r[i] = f[i] - s;
    }
    for (int Z = 1; Z < gc.n - 1; ++Z) for (int Y = 1; Y < gc.n - 1; ++Y) for (int X = 1; X < gc.n - 1; ++X) {
        double acc = 0.0;
        for (int dz = 0; dz < 3; ++dz) for (int dy = 0; dy < 3; ++dy) for (int dx = 0; dx < 3; ++dx)
            acc = acc + W.w[dz * 9 + dy * 3 + dx] * r[(2 * Z + dz - 1) * gf.plane + (long long)(2 * Y + dy - 1) * gf.pitch + 2 * X + dx - 1];
        ref[Z * gc.plane + (long long)Y * gc.pitch + X] = acc;
    }
    double *du, *df, *dd;
    cudaMalloc(&du, gf.total * 8); cudaMalloc(&df, gf.total * 8); cudaMalloc(&dd, gc.total * 8);
    cudaMemcpy(du, u.data(), gf.total * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(df, f.data(), gf.total * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dd, out.data(), gc.total * 8, cudaMemcpyHostToDevice);
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    bool ok = launch_rr_col<NW_, RC_, NPS_, MINB_>(pr.multiProcessorCount, gf, gc, c, W, du, df, dd, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("level %d launch %d sync: %s\n", level, (int)ok, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(out.data(), dd, gc.total * 8, cudaMemcpyDeviceToHost);
    long long bad = 0, cnt = 0;
    for (int Z = 1; Z < gc.n - 1; ++Z) for (int Y = 1; Y < gc.n - 1; ++Y) for (int X = 1; X < gc.n - 1; ++X) {
        const long long i = Z * gc.plane + (long long)Y * gc.pitch + X;
        ++cnt;
        if (out[i] != ref[i]) { if (bad < 10) printf("  mismatch at (%d,%d,%d): %.17g vs %.17g\n", X, Y, Z, out[i], ref[i]); ++bad; }
    }
    printf("mismatches %lld of %lld\n", bad, cnt);
    // timing
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch_rr_col<NW_, RC_, NPS_, MINB_>(pr.multiProcessorCount, gf, gc, c, W, du, df, dd, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch_rr_col<NW_, RC_, NPS_, MINB_>(pr.multiProcessorCount, gf, gc, c, W, du, df, dd, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double nd = (double)(gf.n - 2) * (gf.n - 2) * (gf.n - 2);
    printf("%.4f ms per launch, %.1f GB/s\n", ms / 10, 17.0 * nd / (ms / 10 * 1e-3) / 1e9);
    return bad != 0;
}
