// qtorch_b200/csrc/sliced.inl -- the sliced-amplitude executor (included at the end of engine.cu, after plan.inl).
//
// One index-sliced network = 4^s sub-networks with identical step shapes whose scalar values add up to the network
// value (SURVEY 8e "Slices"; the reduction replaces the reference's serial `f_pVal +=`, /root/reference/src/maxcut.cpp:196).
// This object runs the slices a rank owns WITHOUT ever synchronising with the host inside an amplitude:
//
//   * `lanes` replicas of the per-slice part of the plan (own scratch, own CUDA graphs, own stream).  A rank's slices are
//     dealt over the lanes; two lanes keep two big launches queued at any time, so the persistent tile kernel of one slice
//     starts on every SM the previous slice's kernel has left (its prologue and the other's tail overlap), and the
//     single-CTA grouped launches of one lane run beside the other lane's big steps.
//   * the slice-invariant prefix runs ONCE per amplitude on a stream of its own, into one of two prefix buffer sets that
//     all lanes read; with two amplitudes in flight (begin(i+1) before end(i)) the prefix of amplitude i+1 -- hundreds of
//     latency-bound micro-steps on one SM -- runs beside the slices of amplitude i instead of idling 147 SMs.
//   * slot scalars are accumulated ON THE DEVICE in a fixed order (lane by lane, slot by slot: deterministic), the lane
//     sums meet in one tiny kernel on the ctx stream, followed in-stream by one ncclAllReduce of the complex scalar
//     (no host staging) and a 16-byte device->host copy into a pinned slot that `qtb_read_scalar_end` waits for.
//   * per-slot inputs are staged on a separate upload stream, so re-staging for amplitude i+2 does not queue behind the
//     device work of amplitude i+1.

__global__ void k_acc_scalar(double2 *acc, const double2 *v, int first) {
    if (threadIdx.x == 0) {
        double2 a = first ? make_double2(0.0, 0.0) : *acc;
        const double2 x = *v;
        a.x += x.x; a.y += x.y;
        *acc = a;
    }
}
__global__ void k_sum_scalars(double2 *out, const double2 *in, int n) {
    if (threadIdx.x == 0) {
        double sr = 0.0, si = 0.0;
        for (int i = 0; i < n; i++) { sr += in[i].x; si += in[i].y; }
        *out = make_double2(sr, si);
    }
}

struct qtb_sliced_s {
    // Two amplitudes may be in flight (parity p = amplitude index mod 2).  prefixOwner[p] runs the slice-invariant prefix
    // into prefix buffer set p; lane l has one plan replica per parity, compiled onto prefix set p (shared by all lanes)
    // and onto the lane's own suffix scratch (shared by its two parities: a lane's stream serialises them).
    qtb_plan *prefixOwner[2] = {nullptr, nullptr};
    std::vector<qtb_plan *> lanes[2];
    int nLanes = 0;
    bool hasPrefix = false;
    std::vector<cudaStream_t> streams;         // one per lane
    cudaStream_t prefixStream = nullptr, upStream = nullptr;
    std::vector<cudaEvent_t> laneDone;         // [2 * nLanes]
    cudaEvent_t prefixDone[2] = {nullptr, nullptr};
    cudaEvent_t ampDone[2] = {nullptr, nullptr}; bool ampDoneValid[2] = {false, false};
    cudaEvent_t staged = nullptr; bool stagedValid = false;
    std::vector<uint8_t *> slotDev;            // staged small-input blobs, shared by all lanes
    uint8_t *stageHost = nullptr;
    double2 *laneAcc = nullptr, *total = nullptr;      // [2][8], [2]
    size_t blobBytes = 0;
    unsigned long long amplitudes = 0;
    std::vector<qtb_plan *> all() const {
        std::vector<qtb_plan *> v;
        for (int p = 0; p < 2; p++) { if (prefixOwner[p]) v.push_back(prefixOwner[p]); for (qtb_plan *q : lanes[p]) v.push_back(q); }
        return v;
    }
    const qtb_plan *shape() const { return lanes[0][0]; }
};

static int sliced_destroy_locked(qtb_ctx *ctx, qtb_sliced *sl) {
    cudaSetDevice(ctx->device);
    for (cudaStream_t s : sl->streams) if (s) cudaStreamSynchronize(s);
    if (sl->prefixStream) cudaStreamSynchronize(sl->prefixStream);
    if (sl->upStream) cudaStreamSynchronize(sl->upStream);
    cudaStreamSynchronize(ctx->stream);
    for (cudaStream_t s : sl->streams) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : sl->laneDone) if (e) cudaEventDestroy(e);
    for (int p = 0; p < 2; p++) { if (sl->prefixDone[p]) cudaEventDestroy(sl->prefixDone[p]); if (sl->ampDone[p]) cudaEventDestroy(sl->ampDone[p]); }
    for (uint8_t *p : sl->slotDev) if (p) cudaFree(p);
    if (sl->stageHost) cudaFreeHost(sl->stageHost);
    if (sl->prefixStream) cudaStreamDestroy(sl->prefixStream);
    if (sl->upStream) cudaStreamDestroy(sl->upStream);
    if (sl->staged) cudaEventDestroy(sl->staged);
    if (sl->laneAcc) cudaFree(sl->laneAcc);
    if (sl->total) cudaFree(sl->total);
    cudaGetLastError();
    return QTB_OK;
}

static int sliced_init_locked(qtb_ctx *ctx, qtb_sliced *sl) {
    sl->blobBytes = sl->shape()->inBlobBytes;
    CU(cudaSetDevice(ctx->device));
    for (int l = 0; l < sl->nLanes; l++) {
        cudaStream_t s = nullptr;
        CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        sl->streams.push_back(s);
    }
    for (int i = 0; i < 2 * sl->nLanes; i++) {
        cudaEvent_t e = nullptr;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sl->laneDone.push_back(e);
    }
    for (int p = 0; p < 2; p++) {
        CU(cudaEventCreateWithFlags(&sl->prefixDone[p], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&sl->ampDone[p], cudaEventDisableTiming));
    }
    CU(cudaStreamCreateWithFlags(&sl->prefixStream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&sl->upStream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&sl->staged, cudaEventDisableTiming));
    CU(cudaMalloc((void **)&sl->laneAcc, sizeof(double2) * 16));
    CU(cudaMalloc((void **)&sl->total, sizeof(double2) * 2));
    CU(cudaMallocHost((void **)&sl->stageHost, std::max<size_t>(sl->blobBytes, 256)));
    return QTB_OK;
}

extern "C" {

int qtb_sliced_create(qtb_ctx *ctx, int nInputs, const int *inputRanks, int nSteps, const qtb_plan_step *steps, int nPrefix,
                      int nLanes, qtb_sliced **out) {
    if (!ctx || !out) return fail(QTB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (nLanes < 1 || nLanes > 8) return fail(QTB_ERR_INVALID, "lanes must be 1..8");
    for (int i = 0; i < nInputs; i++)
        if (inputRanks && inputRanks[i] > 5) return fail(QTB_ERR_UNSUPPORTED, "sliced plans need every input to have rank <= 5 (gate / state / measurement tensors)");
    if (nPrefix < 0 || nPrefix > nSteps) return fail(QTB_ERR_INVALID, "bad invariant-prefix length");
    if (nPrefix == nSteps) nPrefix = 0;
    qtb_sliced *sl = new qtb_sliced_s();
    sl->nLanes = nLanes;
    sl->hasPrefix = nPrefix > 0;
    auto bail = [&](int st) {
        const std::string keep = g_lastError;
        for (qtb_plan *p : sl->all()) qtb_plan_destroy(ctx, p);
        { std::lock_guard<std::mutex> lk(ctx->mu); sliced_destroy_locked(ctx, sl); }
        delete sl;
        g_lastError = keep;
        return st;
    };
    for (int par = 0; par < 2; par++) {
        if (sl->hasPrefix) {
            const int st = plan_create_replica(ctx, nInputs, inputRanks, nSteps, steps, nPrefix, nullptr, nullptr, true, &sl->prefixOwner[par]);
            if (st != QTB_OK) return bail(st);
        }
        for (int l = 0; l < nLanes; l++) {
            qtb_plan *p = nullptr;
            const int st = plan_create_replica(ctx, nInputs, inputRanks, nSteps, steps, nPrefix, sl->prefixOwner[par],
                                               par == 1 ? sl->lanes[0][l] : nullptr, false, &p);
            if (st != QTB_OK) return bail(st);
            sl->lanes[par].push_back(p);
            if (p->outRank != 0) return bail(fail(QTB_ERR_INVALID, "a sliced plan must end in a scalar"));
        }
    }
    int st = QTB_OK;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        st = sliced_init_locked(ctx, sl);
    }
    if (st != QTB_OK) return bail(st);
    *out = sl;
    return QTB_OK;
}

int qtb_sliced_destroy(qtb_ctx *ctx, qtb_sliced *sl) {
    if (!sl) return QTB_OK;
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    { std::lock_guard<std::mutex> lk(ctx->mu); sliced_destroy_locked(ctx, sl); }
    for (qtb_plan *p : sl->all()) qtb_plan_destroy(ctx, p);
    delete sl;
    return QTB_OK;
}

int qtb_sliced_stage(qtb_ctx *ctx, qtb_sliced *sl, int slot, const double *const *hostInputs) {
    if (!ctx || !sl || slot < 0 || slot > 65535 || !hostInputs) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    const qtb_plan *p0 = sl->shape();
    if ((size_t)slot >= sl->slotDev.size()) sl->slotDev.resize(slot + 1, nullptr);
    if (!sl->slotDev[slot]) CU(cudaMalloc((void **)&sl->slotDev[slot], std::max<size_t>(sl->blobBytes, 256)));
    if (sl->stagedValid) CU(cudaEventSynchronize(sl->staged));          // the pinned staging buffer is free again
    for (int i = 0; i < p0->nInputs; i++) {
        if (!hostInputs[i]) return fail(QTB_ERR_EMPTY_INPUT, "null input tensor");
        memcpy(sl->stageHost + p0->inputBlobOff[i], hostInputs[i], Pool::bytes(p0->inputRanks[i]));
        ctx->stats.bytes_h2d += (long long)Pool::bytes(p0->inputRanks[i]);
    }
    CU(cudaMemcpyAsync(sl->slotDev[slot], sl->stageHost, sl->blobBytes, cudaMemcpyHostToDevice, sl->upStream));
    CU(cudaEventRecord(sl->staged, sl->upStream));
    sl->stagedValid = true;
    return QTB_OK;
}

int qtb_sliced_begin(qtb_ctx *ctx, qtb_sliced *sl, const int *slots, int n, int allreduce, qtb_scalar_read **out) {
    if (!ctx || !sl || !out || n < 0 || (n > 0 && !slots)) return fail(QTB_ERR_INVALID, "bad argument");
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (int j = 0; j < n; j++)
        if (slots[j] < 0 || (size_t)slots[j] >= sl->slotDev.size() || !sl->slotDev[slots[j]]) return fail(QTB_ERR_INVALID, "unknown input slot");
    if (allreduce && !ctx->comm) return fail(QTB_ERR_NCCL, "communicator not initialised (qtb_comm_init)");
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    const int par = (int)(sl->amplitudes & 1ull);
    sl->amplitudes++;
    const int R = sl->nLanes, used = std::min(R, n);
    const bool useGraph = plan_graphs_enabled() && !ctx->trace;
    // graphs: captured on the ctx stream the first time a replica is used
    if (useGraph) {
        qtb_plan *po = sl->prefixOwner[par];
        if (po && n > 0 && !po->partGraphsTried) { po->partGraphsTried = true; if (po->segs.size() >= 2) po->graphPrefix = plan_capture(ctx, po, 0, po->segs.size()); }
        for (int l = 0; l < used; l++) {
            qtb_plan *pl = sl->lanes[par][l];
            if (!pl->partGraphsTried) { pl->partGraphsTried = true; if (pl->segs.size() - pl->prefixSegs >= 2) pl->graphSuffix = plan_capture(ctx, pl, pl->prefixSegs, pl->segs.size()); }
        }
    }
    const qtb_plan *shape = sl->shape();
    int preSteps = 0, preMicro = 0;
    for (size_t k = 0; k < shape->prefixSegs; k++) { preSteps += shape->segs[k].nSteps; if (shape->segs[k].micro) preMicro += shape->segs[k].nSteps; }
    // ---- the slice-invariant prefix: once per amplitude, on its own stream, into prefix buffer set `par`
    if (sl->hasPrefix && n > 0) {
        qtb_plan *po = sl->prefixOwner[par];
        cudaStream_t ps = sl->prefixStream;
        if (sl->stagedValid) CU(cudaStreamWaitEvent(ps, sl->staged, 0));
        // buffer set `par` was last read by the slices of amplitude i-2: they are folded into that amplitude's total by now?
        if (sl->ampDoneValid[par]) CU(cudaStreamWaitEvent(ps, sl->ampDone[par], 0));
        CU(cudaMemcpyAsync(po->inBlobDev, sl->slotDev[slots[0]], sl->blobBytes, cudaMemcpyDeviceToDevice, ps));
        if (useGraph && po->graphPrefix) { CU(cudaGraphLaunch(po->graphPrefix, ps)); ctx->stats.launches += po->launches; }
        else ST(plan_enqueue(ctx, po, ps, 0, po->segs.size()));
        CU(cudaEventRecord(sl->prefixDone[par], ps));
        ctx->stats.steps += preSteps; ctx->stats.micro_steps += preMicro; ctx->stats.units += shape->prefixUnits;
    }
    // ---- the slices, dealt over the lanes
    for (int l = 0; l < used; l++) {
        qtb_plan *pl = sl->lanes[par][l];
        cudaStream_t s = sl->streams[l];
        const size_t nSegs = pl->segs.size(), pre = pl->prefixSegs;
        if (sl->stagedValid) CU(cudaStreamWaitEvent(s, sl->staged, 0));
        if (sl->hasPrefix) CU(cudaStreamWaitEvent(s, sl->prefixDone[par], 0));
        bool first = true;
        for (int j = l; j < n; j += R) {
            CU(cudaMemcpyAsync(pl->inBlobDev, sl->slotDev[slots[j]], sl->blobBytes, cudaMemcpyDeviceToDevice, s));
            if (useGraph && pl->graphSuffix) { CU(cudaGraphLaunch(pl->graphSuffix, s)); ctx->stats.launches += pl->launches - pl->launchesPrefix; }
            else ST(plan_enqueue(ctx, pl, s, pre, nSegs));
            // accumulator set `par` belongs to amplitude i-2 until that amplitude's total has been formed
            if (first && sl->ampDoneValid[par]) CU(cudaStreamWaitEvent(s, sl->ampDone[par], 0));
            k_acc_scalar<<<1, 32, 0, s>>>(sl->laneAcc + 8 * par + l, pl->outDev, first ? 1 : 0);
            ctx->stats.launches++;
            first = false;
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(sl->laneDone[par * R + l], s));
        CU(cudaStreamWaitEvent(ctx->stream, sl->laneDone[par * R + l], 0));
    }
    ctx->stats.steps += (long long)n * (shape->nSteps - preSteps);
    ctx->stats.micro_steps += (long long)n * (shape->nMicroSteps - preMicro);
    ctx->stats.units += (long long)n * (shape->units - shape->prefixUnits);
    k_sum_scalars<<<1, 32, 0, ctx->stream>>>(sl->total + par, sl->laneAcc + 8 * par, used);       // used == 0: this rank owns no slice, total = 0
    CU(cudaGetLastError());
    ctx->stats.launches++;
    CU(cudaEventRecord(sl->ampDone[par], ctx->stream));
    sl->ampDoneValid[par] = true;
    if (allreduce) {
        int r = g_nccl.AllReduce(sl->total + par, sl->total + par, 2, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
        if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    }
    qtb_scalar_read *rd = nullptr;
    ST(scalar_read_enqueue(ctx, sl->total + par, &rd));
    *out = rd;
    return QTB_OK;
}

int qtb_sliced_lanes(qtb_sliced *sl) { return sl ? sl->nLanes : 0; }
long long qtb_sliced_units(qtb_sliced *sl) { return sl ? sl->shape()->units : 0; }
long long qtb_sliced_prefix_units(qtb_sliced *sl) { return sl ? sl->shape()->prefixUnits : 0; }
int qtb_sliced_launches(qtb_sliced *sl, int *prefixLaunches) {
    if (!sl) return 0;
    if (prefixLaunches) *prefixLaunches = sl->shape()->launchesPrefix;
    return sl->shape()->launches;
}

}  // extern "C"
