// Stand-alone bring-up test of the tcgen05 fp16-split GEMMs (not part of libcpros.so).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o build/tc_gemm_test <this file>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../gemm_tc.cuh"

unsigned long long g_cp_launches = 0;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int run_nt(int64_t M, int N, int K, bool check, int reps) {
    std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N);
    for (auto& v : A) v = frand();
    for (auto& v : B) v = frand() * 0.1f;
    for (auto& v : bias) v = frand();
    float *dA, *dB, *dbias, *dC, *dps, *dpq;
    plane_t *dAh, *dAl, *dBh, *dBl;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dAh, A.size() * 2)); CK(cudaMalloc(&dAl, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 2)); CK(cudaMalloc(&dBl, B.size() * 2));
    CK(cudaMalloc(&dbias, N * 4)); CK(cudaMalloc(&dC, (size_t)M * N * 4));
    const int64_t tm = (M + 127) / 128;
    CK(cudaMalloc(&dps, tm * N * 4)); CK(cudaMalloc(&dpq, tm * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dbias, bias.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dC, 0xff, (size_t)M * N * 4));
    split_planes_kernel<<<1024, 256>>>(dA, dAh, dAl, (int64_t)A.size() / 4, 1.f);
    split_planes_kernel<<<256, 256>>>(dB, dBh, dBl, (int64_t)B.size() / 4, 1.f);
    CK(cudaDeviceSynchronize());
    int rc = tcg::launch_nt(dAh, dAl, M, K, K, dBh, dBl, N, K, dbias, dC, N, dps, dpq, 1, 0);
    if (rc) { printf("launch rc=%d\n", rc); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    int bad = 0;
    if (check) {
        std::vector<float> C((size_t)M * N), ps(tm * N), pq(tm * N);
        CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ps.data(), dps, ps.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pq.data(), dpq, pq.size() * 4, cudaMemcpyDeviceToHost));
        double num = 0, den = 0, maxabs = 0;
        std::vector<double> cs(N, 0.0), cq(N, 0.0);
        for (int64_t i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                double s = bias[j];
                for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * (double)B[(size_t)j * K + k];
                if (s < 0) s = 0;
                const double d = (double)C[i * N + j] - s;
                num += d * d; den += s * s;
                if (fabs(d) > maxabs) maxabs = fabs(d);
                cs[j] += s; cq[j] += s * s;
            }
        double snum = 0, sden = 0, qnum = 0, qden = 0;
        for (int j = 0; j < N; ++j) {
            double s = 0, q = 0;
            for (int64_t t = 0; t < tm; ++t) { s += ps[t * N + j]; q += pq[t * N + j]; }
            snum += (s - cs[j]) * (s - cs[j]); sden += cs[j] * cs[j];
            qnum += (q - cq[j]) * (q - cq[j]); qden += cq[j] * cq[j];
        }
        const double rel = sqrt(num / den);
        printf("NT M=%lld N=%d K=%d: rel err %.3e  max abs %.3e  colsum rel %.3e  colsq rel %.3e\n", (long long)M, N, K,
               rel, maxabs, sqrt(snum / sden), sqrt(qnum / qden));
        bad = !(rel < 5e-6);
    }
    if (reps > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int variant = 0; variant < 2; ++variant) {
            float* ps = variant == 0 ? dps : nullptr;
            float* pq = variant == 0 ? dpq : nullptr;
            for (int i = 0; i < 3; ++i) tcg::launch_nt(dAh, dAl, M, K, K, dBh, dBl, N, K, dbias, dC, N, ps, pq, 1, 0);
            cudaEventRecord(e0);
            for (int i = 0; i < reps; ++i) tcg::launch_nt(dAh, dAl, M, K, K, dBh, dBl, N, K, dbias, dC, N, ps, pq, 1, 0);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
            printf("NT M=%lld N=%d K=%d stats=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent), %.1f TFLOP/s fp16 issued\n",
                   (long long)M, N, K, variant == 0, ms, 2.0 * M * N * K / ms / 1e9, 6.0 * M * N * K / ms / 1e9);
        }
    }
    cudaFree(dA); cudaFree(dAh); cudaFree(dAl); cudaFree(dB); cudaFree(dBh); cudaFree(dBl); cudaFree(dbias); cudaFree(dC);
    cudaFree(dps); cudaFree(dpq);
    return bad;
}

int run_tn(int64_t R, int Mo, int No, bool check, int reps) {
    std::vector<float> G((size_t)R * Mo), A((size_t)R * No);
    for (auto& v : G) v = frand();
    for (auto& v : A) v = frand() + 0.3f;
    float *dG, *dA, *dP;
    plane_t *dGh, *dGl, *dAh, *dAl;
    const size_t cap = (size_t)32 * Mo * No;
    CK(cudaMalloc(&dG, G.size() * 4)); CK(cudaMalloc(&dGh, G.size() * 2)); CK(cudaMalloc(&dGl, G.size() * 2));
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dAh, A.size() * 2)); CK(cudaMalloc(&dAl, A.size() * 2));
    CK(cudaMalloc(&dP, cap * 4));
    CK(cudaMemcpy(dG, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    split_planes_kernel<<<1024, 256>>>(dG, dGh, dGl, (int64_t)G.size() / 4, 1.f);
    split_planes_kernel<<<1024, 256>>>(dA, dAh, dAl, (int64_t)A.size() / 4, 1.f);
    CK(cudaDeviceSynchronize());
    int S = 0;
    int rc = tcg::launch_tn(dGh, dGl, Mo, Mo, dAh, dAl, No, No, R, dP, cap, &S, 0);
    if (rc) { printf("launch_tn rc=%d\n", rc); return 1; }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tn kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    int bad = 0;
    if (check) {
        std::vector<float> P((size_t)S * Mo * No);
        CK(cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost));
        double num = 0, den = 0;
        for (int o = 0; o < Mo; o += 7)
            for (int c = 0; c < No; c += 5) {
                double s = 0;
                for (int64_t r = 0; r < R; ++r) s += (double)G[r * Mo + o] * (double)A[r * No + c];
                double got = 0;
                for (int z = 0; z < S; ++z) got += P[((size_t)z * Mo + o) * No + c];
                num += (got - s) * (got - s); den += s * s;
            }
        const double rel = sqrt(num / den);
        printf("TN R=%lld Mo=%d No=%d splits=%d: rel err %.3e\n", (long long)R, Mo, No, S, rel);
        bad = !(rel < 5e-6);
    }
    if (reps > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) tcg::launch_tn(dGh, dGl, Mo, Mo, dAh, dAl, No, No, R, dP, cap, &S, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) tcg::launch_tn(dGh, dGl, Mo, Mo, dAh, dAl, No, No, R, dP, cap, &S, 0);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("TN R=%lld Mo=%d No=%d splits=%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent)\n", (long long)R, Mo, No, S, ms,
               2.0 * R * Mo * No / ms / 1e9);
    }
    cudaFree(dG); cudaFree(dGh); cudaFree(dGl); cudaFree(dA); cudaFree(dAh); cudaFree(dAl); cudaFree(dP);
    return bad;
}

int main(int argc, char** argv) {
    int bad = 0;
    if (argc > 1 && argv[1][0] == '1') tcg::g_use_pair = false;
    printf("pair kernel: %d\n", (int)tcg::g_use_pair);
    const bool quick = argc > 2;
    if (!quick) {
        bad |= run_nt(128, 128, 64, true, 0);
        bad |= run_nt(128, 128, 512, true, 0);
        bad |= run_nt(1000, 512, 768, true, 0);
        bad |= run_nt(4100, 512, 512, true, 0);
        bad |= run_tn(64, 128, 128, true, 0);
        bad |= run_tn(48, 128, 128, true, 0);
        bad |= run_tn(1000, 128, 128, true, 0);
        bad |= run_tn(5000, 512, 768, true, 0);
        bad |= run_tn(41 * 1000, 512, 512, true, 0);
    }
    if (!bad) run_nt(167936, 512, 512, false, 10);
    if (!bad) run_nt(167936, 768, 512, false, 10);
    if (!bad) run_tn(167936, 512, 512, false, 10);
    if (!bad) run_tn(167936, 512, 768, false, 10);
    printf(bad ? "FAILED\n" : "OK\n");
    return bad;
}
