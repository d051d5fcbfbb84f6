// tools/probe_tma.cu -- does TMA (cp.async.bulk.tensor) gather a base-4 digit tile of a Node tensor?
//
// A rank-r tensor is 4^r complex<double> (16 B), leg 0 fastest.  A tile = a few legs taken whole (runs of contiguous
// legs) at an arbitrary element offset (the other legs' digits).  Tensor map used by the engine's TMA path:
//   dim0 = (re, im)                       2 x f64, contiguous
//   dim1 = "offset" dimension             size = 4^r elements, stride 16 B, box extent 1  -> coordinate = element offset
//   dim2.. = one dimension per run        size 4^len, stride 16 * 4^pos bytes, box extent = size
// The dimensions alias each other in memory; this probe checks that the driver accepts such a map, that the copy
// returns the right elements, and what the shared-memory layout is with and without the 128-byte swizzle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_build/probe_tma tools/probe_tma.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_probe(const __grid_constant__ CUtensorMap tm, int nd, int off, double2 *out, int boxElems) {
    extern __shared__ uint8_t smemRaw[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smemRaw) + 1023) & ~(uintptr_t)1023);   // swizzled boxes want 1024 B
    const uint32_t barAddr = (uint32_t)__cvta_generic_to_shared(&bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = threadIdx.x; i < 3072; i += blockDim.x) reinterpret_cast<double2 *>(smem)[i] = make_double2(-7.0, -7.0);   // sentinel
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(boxElems * 16) : "memory");
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                         ::"r"(dst), "l"(&tm), "r"(0), "r"(off), "r"(0), "r"(0), "r"(0), "r"(barAddr) : "memory");
    }
    // bounded wait: a wrong byte count must not hang the box
    bool done = false;
    for (int spin = 0; spin < 2000000 && !done; spin++) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(barAddr), "r"(0) : "memory");
        done = ok != 0;
    }
    if (threadIdx.x == 0 && !done) out[boxElems] = make_double2(-1.0, -1.0);
    __syncthreads();
    const double2 *s = reinterpret_cast<const double2 *>(smem);
    for (int i = threadIdx.x; i < boxElems; i += blockDim.x) out[i] = s[i];
    for (int i = threadIdx.x; i < 3072; i += blockDim.x) out[2048 + i] = s[i];            // the whole window, to see where things landed
}

int main() {
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) { printf("{\"probe\":\"tma\",\"error\":\"no cuTensorMapEncodeTiled\"}\n"); return 1; }
    const int rank = 10;
    const size_t n = (size_t)1 << (2 * rank);
    std::vector<double2> h(n);
    for (size_t i = 0; i < n; i++) h[i] = make_double2((double)i, -(double)i);
    double2 *d = nullptr, *out = nullptr;
    cudaMalloc(&d, n * 16);
    cudaMemcpy(d, h.data(), n * 16, cudaMemcpyHostToDevice);
    cudaMalloc(&out, (2048 + 3072) * 16);
    struct Case { const char *name; int nruns; int pos[3], len[3]; CUtensorMapSwizzle sw; };
    const Case cases[] = {
        {"one run legs 0-3, no swizzle", 1, {0, 0, 0}, {4, 0, 0}, CU_TENSOR_MAP_SWIZZLE_NONE},
        {"runs legs 0-1 and 4-5, no swizzle", 2, {0, 4, 0}, {2, 2, 0}, CU_TENSOR_MAP_SWIZZLE_NONE},
        {"runs legs 0-1, 3, 6-7, no swizzle", 3, {0, 3, 6}, {2, 1, 2}, CU_TENSOR_MAP_SWIZZLE_NONE},
        {"one run legs 0-3, swizzle 32B", 1, {0, 0, 0}, {4, 0, 0}, CU_TENSOR_MAP_SWIZZLE_32B},
        {"one run legs 0-3, swizzle 64B", 1, {0, 0, 0}, {4, 0, 0}, CU_TENSOR_MAP_SWIZZLE_64B},
        {"one run legs 0-3, swizzle 128B", 1, {0, 0, 0}, {4, 0, 0}, CU_TENSOR_MAP_SWIZZLE_128B},
        {"runs legs 0-1 and 4-5, swizzle 128B", 2, {0, 4, 0}, {2, 2, 0}, CU_TENSOR_MAP_SWIZZLE_128B},
        {"runs legs 0-1, 3, 6-7, swizzle 128B", 3, {0, 3, 6}, {2, 1, 2}, CU_TENSOR_MAP_SWIZZLE_128B},
        {"runs leg 2 and legs 5-6 (tile without leg 0), swizzle 128B", 2, {2, 5, 0}, {1, 2, 0}, CU_TENSOR_MAP_SWIZZLE_128B},
    };
    // ---- the engine's layout: dim0 = leg 0 x (re, im) = 64 bytes, offset dimension in units of 4 elements, 64-byte swizzle
    {
        struct Case2 { const char *name; int nruns; int pos[3], len[3]; };
        const Case2 cases2[] = {
            {"leg 0 + run legs 1-4, swizzle 64B", 1, {1, 0, 0}, {4, 0, 0}},
            {"leg 0 + runs leg 2, legs 4-5, leg 9, swizzle 64B", 3, {2, 4, 9}, {1, 2, 1}},
            {"leg 0 + run legs 6-9, swizzle 64B", 1, {6, 0, 0}, {4, 0, 0}},
        };
        for (const Case2 &c : cases2) {
            cuuint64_t gdim[5] = {8, (cuuint64_t)n / 4, 1, 1, 1};
            cuuint64_t gstr[4] = {64, (cuuint64_t)16 << (2 * rank), (cuuint64_t)16 << (2 * rank), (cuuint64_t)16 << (2 * rank)};
            cuuint32_t box[5] = {8, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
            int boxElems = 4;
            for (int r = 0; r < c.nruns; r++) {
                gdim[2 + r] = (cuuint64_t)1 << (2 * c.len[r]);
                gstr[1 + r] = (cuuint64_t)16 << (2 * c.pos[r]);
                box[2 + r] = (cuuint32_t)gdim[2 + r];
                boxElems *= (int)gdim[2 + r];
            }
            CUtensorMap tm;
            memset(&tm, 0, sizeof(tm));
            CUresult res = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (res != CUDA_SUCCESS) { printf("{\"probe\":\"tma\",\"case\":\"%s\",\"encode\":%d}\n", c.name, (int)res); continue; }
            // corner: digits on legs that are in no run (leg 1 or 3 where free, leg 8 ...), leg-0 digit 0
            int off = 0;
            auto inRun = [&](int leg) { for (int r = 0; r < c.nruns; r++) if (c.pos[r] <= leg && leg < c.pos[r] + c.len[r]) return true; return leg == 0; };
            for (int leg = 1; leg < rank; leg++) if (!inRun(leg)) off |= ((leg % 3) + 1) << (2 * leg);
            cudaMemset(out, 0, (2048 + 3072) * 16);
            cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * 16 + 1024);
            k_probe<<<1, 128, 3072 * 16 + 1024>>>(tm, 5, off / 4, out, boxElems);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<double2> got(boxElems + 1);
            cudaMemcpy(got.data(), out, (boxElems + 1) * 16, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int idx = 0; idx < boxElems; idx++) {
                size_t el = (size_t)off + (idx & 3);
                int rem = idx >> 2;
                for (int r = 0; r < c.nruns; r++) { const int sz = 1 << (2 * c.len[r]); el += (size_t)(rem % sz) << (2 * c.pos[r]); rem /= sz; }
                const int slot = idx ^ ((idx >> 3) & 3);
                if (got[slot].x != (double)el) bad++;
            }
            printf("{\"probe\":\"tma\",\"case\":\"%s\",\"encode\":0,\"cuda\":\"%s\",\"timeout\":%d,\"box_elems\":%d,\"mismatch_slot_eq_idx_xor_idx_shr3_and3\":%d}\n",
                   c.name, cudaGetErrorString(e), got[boxElems].x == -1.0 ? 1 : 0, boxElems, bad);
        }
    }
    for (const Case &c : cases) {
        const int nd = 5;                                      // always five dimensions: unused ones have size 1 (as the engine does)
        cuuint64_t gdim[5] = {2, (cuuint64_t)n, 1, 1, 1};
        cuuint64_t gstr[4] = {16, (cuuint64_t)16 << (2 * rank), (cuuint64_t)16 << (2 * rank), (cuuint64_t)16 << (2 * rank)};   // strides of dims 1..4 in bytes
        cuuint32_t box[5] = {2, 1, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
        int boxElems = 1;
        for (int r = 0; r < c.nruns; r++) {
            gdim[2 + r] = (cuuint64_t)1 << (2 * c.len[r]);
            gstr[1 + r] = (cuuint64_t)16 << (2 * c.pos[r]);
            box[2 + r] = (cuuint32_t)gdim[2 + r];
            boxElems *= (int)gdim[2 + r];
        }
        CUtensorMap tm;
        memset(&tm, 0, sizeof(tm));
        CUresult res = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, nd, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (res != CUDA_SUCCESS) { printf("{\"probe\":\"tma\",\"case\":\"%s\",\"encode\":%d}\n", c.name, (int)res); continue; }
        // an offset that sets digits of legs outside the runs (legs 8, 9 here, and leg 2 where it is free)
        int off = (2 << 16) | (1 << 18);
        bool usesLeg2 = false;
        for (int r = 0; r < c.nruns; r++) if (c.pos[r] <= 2 && 2 < c.pos[r] + c.len[r]) usesLeg2 = true;
        if (!usesLeg2) off |= 3 << 4;
        cudaMemset(out, 0, (2048 + 3072) * 16);
        cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 3072 * 16 + 1024);
        k_probe<<<1, 128, 3072 * 16 + 1024>>>(tm, nd, off, out, boxElems);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<double2> got(boxElems + 1);
        cudaMemcpy(got.data(), out, (boxElems + 1) * 16, cudaMemcpyDeviceToHost);
        // expected dense order: run 0 fastest, then run 1, ...
        int bad_dense = 0, bad_swz = 0;
        for (int idx = 0; idx < boxElems; idx++) {
            size_t el = (size_t)off;
            int rem = idx;
            for (int r = 0; r < c.nruns; r++) {
                const int sz = 1 << (2 * c.len[r]);
                el += (size_t)(rem % sz) << (2 * c.pos[r]);
                rem /= sz;
            }
            const int swz = idx ^ ((idx >> 3) & 7);
            if (got[idx].x != (double)el) bad_dense++;
            if (got[swz].x != (double)el) bad_swz++;
        }
        {
            std::vector<double2> win(3072);
            cudaMemcpy(win.data(), out + 2048, 3072 * 16, cudaMemcpyDeviceToHost);
            printf("   where the first 24 box elements landed (16-byte slot numbers):");
            for (int idx = 0; idx < 24 && idx < boxElems; idx++) {
                size_t el = (size_t)off; int rem = idx;
                for (int r = 0; r < c.nruns; r++) { const int sz = 1 << (2 * c.len[r]); el += (size_t)(rem % sz) << (2 * c.pos[r]); rem /= sz; }
                int found = -1, count = 0;
                for (int sl = 0; sl < 3072; sl++) if (win[sl].x == (double)el) { if (found < 0) found = sl; count++; }
                printf(" %d", found);
                (void)count;
            }
            int written = 0, last = -1;
            for (int sl = 0; sl < 3072; sl++) if (win[sl].x != -7.0) { written++; last = sl; }
            printf("  | slots written %d, last %d\n", written, last);
        }
        printf("{\"probe\":\"tma\",\"case\":\"%s\",\"encode\":0,\"cuda\":\"%s\",\"timeout\":%d,\"box_elems\":%d,\"mismatch_dense_layout\":%d,\"mismatch_swizzle128_layout\":%d}\n",
               c.name, cudaGetErrorString(e), got[boxElems].x == -1.0 ? 1 : 0, boxElems, bad_dense, bad_swz);
    }
    return 0;
}
