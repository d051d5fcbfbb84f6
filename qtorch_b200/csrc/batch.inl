// qtorch_b200/csrc/batch.inl -- term batches: n independent compiled plans evaluated as ONE CUDA graph per call
// (included at the end of engine.cu).
//
// The QAOA objective F_p = sum_edges 1/2 (1 - Re<Z_i Z_j>) (/root/reference/src/maxcut.cpp:162-204) is evaluated hundreds
// of times by the optimiser with the SAME per-edge networks and only 2p angles changing.  A batch keeps every plan's
// inputs resident in HBM; per evaluation the host hands over the 2p gate tables (Rz(-gamma_l), Rx(2 beta_l): 256 bytes
// each, built by the host mirror exactly as Node.h:271-294 builds them), and one graph launch does
//     H2D of the tables -> scatter into every (plan, input) bound to a table -> all plans -> gather of the n scalars and
//     their sum -> [in-stream ncclAllReduce of the sum] -> one D2H of (sum, n scalars)
// Plans that consist of grouped micro-steps only share one launch (one CTA per plan); others are forked over side streams
// inside the captured graph.

__global__ void k_scatter_tables(const uint64_t *__restrict__ dst, const uint32_t *__restrict__ which, const double2 *__restrict__ tables,
                                 uint32_t nBind, uint32_t elems) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nBind * elems) return;
    const uint32_t b = t / elems, e = t - b * elems;
    reinterpret_cast<double2 *>(dst[b])[e] = tables[(size_t)which[b] * elems + e];
}
// results[0] = sum of the n plan outputs (fixed order: deterministic), results[1 + i] = output of plan i
__global__ void k_gather_terms(const uint64_t *__restrict__ outPtrs, uint32_t n, double2 *results) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) results[1 + i] = *reinterpret_cast<const double2 *>(outPtrs[i]);
    __syncthreads();
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
        for (uint32_t i = 0; i < n; i++) { sr += results[1 + i].x; si += results[1 + i].y; }
        results[0] = make_double2(sr, si);
    }
}

struct qtb_batch_s {
    std::vector<qtb_plan *> plans;
    bool allMicro = true;
    int nTables = 0, tableRank = 2; uint32_t tableElems = 16;
    std::vector<uint64_t> bindDst; std::vector<uint32_t> bindTable; bool bindDirty = true;
    uint64_t *bindDstDev = nullptr; uint32_t *bindTableDev = nullptr; size_t bindCap = 0;
    double2 *tablesDev = nullptr; double *tablesHost = nullptr;
    uint64_t *blobAddrDev = nullptr, *outPtrsDev = nullptr;
    double2 *resultsDev = nullptr; double *resultsHost = nullptr;
    cudaGraphExec_t graph = nullptr; bool graphTried = false;
    cudaEvent_t done = nullptr; bool inFlight = false;
    long long units = 0; int nSteps = 0, nMicro = 0, launches = 0;
};

static void batch_free(qtb_batch *b) {
    if (b->graph) cudaGraphExecDestroy(b->graph);
    if (b->bindDstDev) cudaFree(b->bindDstDev);
    if (b->bindTableDev) cudaFree(b->bindTableDev);
    if (b->tablesDev) cudaFree(b->tablesDev);
    if (b->tablesHost) cudaFreeHost(b->tablesHost);
    if (b->blobAddrDev) cudaFree(b->blobAddrDev);
    if (b->outPtrsDev) cudaFree(b->outPtrsDev);
    if (b->resultsDev) cudaFree(b->resultsDev);
    if (b->resultsHost) cudaFreeHost(b->resultsHost);
    if (b->done) cudaEventDestroy(b->done);
    cudaGetLastError();
    delete b;
}

static int batch_init_locked(qtb_ctx *ctx, qtb_batch *b) {
    const size_t n = b->plans.size();
    CU(cudaMalloc((void **)&b->tablesDev, std::max<size_t>(1, b->nTables) * b->tableElems * sizeof(double2)));
    CU(cudaMallocHost((void **)&b->tablesHost, std::max<size_t>(1, b->nTables) * b->tableElems * sizeof(double2)));
    CU(cudaMalloc((void **)&b->blobAddrDev, n * 8));
    CU(cudaMalloc((void **)&b->outPtrsDev, n * 8));
    CU(cudaMalloc((void **)&b->resultsDev, (n + 1) * sizeof(double2)));
    CU(cudaMallocHost((void **)&b->resultsHost, (n + 1) * sizeof(double2)));
    CU(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
    std::vector<uint64_t> blobs(n), outs(n);
    for (size_t i = 0; i < n; i++) { blobs[i] = reinterpret_cast<uint64_t>(b->plans[i]->microBlobDev); outs[i] = reinterpret_cast<uint64_t>(b->plans[i]->outDev); }
    CU(cudaMemcpy(b->blobAddrDev, blobs.data(), n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b->outPtrsDev, outs.data(), n * 8, cudaMemcpyHostToDevice));
    return QTB_OK;
}

// everything one evaluation enqueues between the table upload and the result copy (captured into the batch's graph)
static int batch_enqueue(qtb_ctx *ctx, qtb_batch *b, bool withReadback) {
    cudaStream_t s = ctx->stream;
    const uint32_t nBind = (uint32_t)b->bindDst.size();
    if (b->nTables > 0) CU(cudaMemcpyAsync(b->tablesDev, b->tablesHost, (size_t)b->nTables * b->tableElems * sizeof(double2), cudaMemcpyHostToDevice, s));
    if (nBind) {
        k_scatter_tables<<<(nBind * b->tableElems + 255) / 256, 256, 0, s>>>(b->bindDstDev, b->bindTableDev, b->tablesDev, nBind, b->tableElems);
        CU(cudaGetLastError());
    }
    const int n = (int)b->plans.size();
    if (b->allMicro) {
        ST(launch_micro_plans(ctx, n, b->units, b->blobAddrDev, s));
    } else {
        const int nAux = std::min(n, 16);
        while ((int)ctx->auxStreams.size() < nAux) {
            cudaStream_t a = nullptr; cudaEvent_t e = nullptr;
            CU(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
            ctx->auxStreams.push_back(a);
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->auxEvents.push_back(e);
        }
        if (!ctx->forkEvent) CU(cudaEventCreateWithFlags(&ctx->forkEvent, cudaEventDisableTiming));
        CU(cudaEventRecord(ctx->forkEvent, s));
        for (int a = 0; a < nAux; a++) CU(cudaStreamWaitEvent(ctx->auxStreams[a], ctx->forkEvent, 0));
        int st = QTB_OK;
        for (int i = 0; i < n && st == QTB_OK; i++) st = plan_enqueue(ctx, b->plans[i], ctx->auxStreams[i % nAux]);
        for (int a = 0; a < nAux; a++) {
            CU(cudaEventRecord(ctx->auxEvents[a], ctx->auxStreams[a]));
            CU(cudaStreamWaitEvent(s, ctx->auxEvents[a], 0));
        }
        ST(st);
    }
    k_gather_terms<<<1, 128, 0, s>>>(b->outPtrsDev, (uint32_t)n, b->resultsDev);
    CU(cudaGetLastError());
    if (withReadback) CU(cudaMemcpyAsync(b->resultsHost, b->resultsDev, (size_t)(n + 1) * sizeof(double2), cudaMemcpyDeviceToHost, s));
    return QTB_OK;
}

extern "C" {

int qtb_batch_create(qtb_ctx *ctx, qtb_plan *const *plans, int n, int nTables, int tableRank, qtb_batch **out) {
    if (!ctx || !plans || !out || n < 1 || nTables < 0 || tableRank < 0 || tableRank > 5) return fail(QTB_ERR_INVALID, "bad argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    qtb_batch *b = new qtb_batch_s();
    b->nTables = nTables; b->tableRank = tableRank; b->tableElems = 1u << (2 * tableRank);
    for (int i = 0; i < n; i++) {
        if (!plans[i] || plans[i]->outRank != 0) { batch_free(b); return fail(QTB_ERR_INVALID, "batched plans must have scalar outputs"); }
        for (int j = 0; j < i; j++) if (plans[j] == plans[i]) { batch_free(b); return fail(QTB_ERR_INVALID, "the same plan appears twice in one batch"); }
        b->plans.push_back(plans[i]);
        if (plans[i]->segs.size() != 1 || !plans[i]->segs[0].micro) b->allMicro = false;
        b->units += plans[i]->units; b->nSteps += plans[i]->nSteps; b->nMicro += plans[i]->nMicroSteps;
    }
    b->launches = 2;                                   // scatter + gather
    if (b->allMicro) b->launches += 1; else for (qtb_plan *p : b->plans) b->launches += p->launches;
    const int st = batch_init_locked(ctx, b);
    if (st != QTB_OK) { const std::string keep = g_lastError; batch_free(b); g_lastError = keep; return st; }
    *out = b;
    return QTB_OK;
}

int qtb_batch_destroy(qtb_ctx *ctx, qtb_batch *b) {
    if (!b) return QTB_OK;
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    batch_free(b);
    return QTB_OK;
}

// all inputs of plan `plan` (host -> the plan's resident buffers); call once per plan before the first run
int qtb_batch_set_inputs(qtb_ctx *ctx, qtb_batch *b, int plan, const double *const *hostInputs) {
    if (!ctx || !b || plan < 0 || plan >= (int)b->plans.size()) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return plan_upload_locked(ctx, b->plans[plan], hostInputs);
}

int qtb_batch_bind(qtb_ctx *ctx, qtb_batch *b, int plan, int input, int table) {
    if (!ctx || !b || plan < 0 || plan >= (int)b->plans.size() || table < 0 || table >= b->nTables) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const qtb_plan *p = b->plans[plan];
    if (input < 0 || input >= p->nInputs || p->inputRanks[input] != b->tableRank) return fail(QTB_ERR_INVALID, "bound input must have the table rank");
    b->bindDst.push_back(reinterpret_cast<uint64_t>(p->inputDev[input]));
    b->bindTable.push_back((uint32_t)table);
    b->bindDirty = true;
    return QTB_OK;
}

int qtb_batch_begin(qtb_ctx *ctx, qtb_batch *b, const double *tables, int allreduce) {
    if (!ctx || !b || (b->nTables > 0 && !tables)) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (allreduce && !ctx->comm) return fail(QTB_ERR_NCCL, "communicator not initialised (qtb_comm_init)");
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    if (b->inFlight) { CU(cudaEventSynchronize(b->done)); b->inFlight = false; }        // pinned staging is free again
    if (b->bindDirty) {
        const size_t nb = b->bindDst.size();
        if (nb > b->bindCap) {
            if (b->bindDstDev) cudaFree(b->bindDstDev);
            if (b->bindTableDev) cudaFree(b->bindTableDev);
            CU(cudaMalloc((void **)&b->bindDstDev, nb * 8));
            CU(cudaMalloc((void **)&b->bindTableDev, nb * 4));
            b->bindCap = nb;
        }
        if (nb) {
            CU(cudaMemcpy(b->bindDstDev, b->bindDst.data(), nb * 8, cudaMemcpyHostToDevice));
            CU(cudaMemcpy(b->bindTableDev, b->bindTable.data(), nb * 4, cudaMemcpyHostToDevice));
        }
        b->bindDirty = false;
        if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
        b->graphTried = false;
    }
    if (b->nTables > 0) memcpy(b->tablesHost, tables, (size_t)b->nTables * b->tableElems * sizeof(double2));
    const bool readInGraph = !allreduce;
    const bool useGraph = plan_graphs_enabled() && !ctx->trace;
    if (useGraph && !b->graphTried) {
        b->graphTried = true;
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const long long keep = ctx->stats.launches;
            const int st = batch_enqueue(ctx, b, readInGraph);
            const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
            ctx->stats.launches = keep;
            if (st == QTB_OK && e == cudaSuccess && g) { if (cudaGraphInstantiate(&b->graph, g, 0) != cudaSuccess) { b->graph = nullptr; cudaGetLastError(); } }
            else cudaGetLastError();
            if (g) cudaGraphDestroy(g);
        } else cudaGetLastError();
    }
    if (useGraph && b->graph) CU(cudaGraphLaunch(b->graph, ctx->stream));
    else ST(batch_enqueue(ctx, b, readInGraph));
    if (allreduce) {
        int r = g_nccl.AllReduce(b->resultsDev, b->resultsDev, 2, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
        if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
        CU(cudaMemcpyAsync(b->resultsHost, b->resultsDev, (b->plans.size() + 1) * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaEventRecord(b->done, ctx->stream));
    b->inFlight = true;
    ctx->stats.launches += b->launches;
    ctx->stats.steps += b->nSteps; ctx->stats.micro_steps += b->nMicro; ctx->stats.units += b->units;
    ctx->stats.bytes_h2d += (long long)b->nTables * b->tableElems * 16;
    ctx->stats.bytes_d2h += (long long)(b->plans.size() + 1) * 16;
    return QTB_OK;
}

// waits for the evaluation started by qtb_batch_begin: sum (over this rank's plans, and over all ranks after an allreduce)
// and, if terms != NULL, this rank's n individual (re, im) pairs
int qtb_batch_end(qtb_ctx *ctx, qtb_batch *b, double sum[2], double *terms) {
    if (!ctx || !b || !sum) return fail(QTB_ERR_INVALID, "bad argument");
    if (!b->inFlight) return fail(QTB_ERR_INVALID, "no evaluation in flight");
    cudaError_t e = cudaEventSynchronize(b->done);
    std::lock_guard<std::mutex> lk(ctx->mu);
    b->inFlight = false;
    if (e != cudaSuccess) return fail(QTB_ERR_CUDA, cudaGetErrorString(e));
    sum[0] = b->resultsHost[0]; sum[1] = b->resultsHost[1];
    if (terms) memcpy(terms, b->resultsHost + 2, b->plans.size() * 16);
    return QTB_OK;
}

int qtb_batch_run(qtb_ctx *ctx, qtb_batch *b, const double *tables, int allreduce, double sum[2], double *terms) {
    ST(qtb_batch_begin(ctx, b, tables, allreduce));
    return qtb_batch_end(ctx, b, sum, terms);
}

int qtb_batch_launches(qtb_batch *b) { return b ? b->launches : 0; }
long long qtb_batch_units(qtb_batch *b) { return b ? b->units : 0; }

}  // extern "C"
