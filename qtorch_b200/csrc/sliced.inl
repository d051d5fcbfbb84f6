// qtorch_b200/csrc/sliced.inl -- the sliced-amplitude executor (included at the end of engine.cu, after plan.inl).
//
// One index-sliced network = 4^s sub-networks with identical step shapes whose scalar values add up to the network
// value (SURVEY 8e "Slices"; the reduction replaces the reference's serial `f_pVal +=`, /root/reference/src/maxcut.cpp:196).
// This object runs the slices a rank owns WITHOUT ever synchronising with the host inside an amplitude:
//
//   * `lanes` replicas of the compiled sliced plan (own buffers, own CUDA graphs, own stream).  A rank's slices are dealt
//     over the lanes; two lanes keep two big launches queued at any time, so the persistent tile kernel of one slice
//     starts on every SM the previous slice's kernel has left (its prologue and the other's tail overlap), and the
//     single-CTA grouped launches of one lane run beside the other lane's big steps.
//   * the slice-invariant prefix runs once per lane and amplitude on the lane's stream; with two amplitudes in flight
//     (begin(i+1) before end(i)) it overlaps the previous amplitude's slices instead of idling 147 SMs.
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
    std::vector<qtb_plan *> lanes;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> laneDone;
    std::vector<uint8_t *> slotDev;            // staged small-input blobs, shared by all lanes
    uint8_t *stageHost = nullptr;
    cudaStream_t upStream = nullptr;
    cudaEvent_t staged = nullptr; bool stagedValid = false;
    cudaEvent_t sumDone = nullptr; bool sumDoneValid = false;
    double2 *laneAcc = nullptr, *total = nullptr;
    size_t blobBytes = 0;
};

static int sliced_destroy_locked(qtb_ctx *ctx, qtb_sliced *sl) {
    cudaSetDevice(ctx->device);
    for (cudaStream_t s : sl->streams) if (s) cudaStreamSynchronize(s);
    if (sl->upStream) cudaStreamSynchronize(sl->upStream);
    cudaStreamSynchronize(ctx->stream);
    for (cudaStream_t s : sl->streams) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : sl->laneDone) if (e) cudaEventDestroy(e);
    for (uint8_t *p : sl->slotDev) if (p) cudaFree(p);
    if (sl->stageHost) cudaFreeHost(sl->stageHost);
    if (sl->upStream) cudaStreamDestroy(sl->upStream);
    if (sl->staged) cudaEventDestroy(sl->staged);
    if (sl->sumDone) cudaEventDestroy(sl->sumDone);
    if (sl->laneAcc) cudaFree(sl->laneAcc);
    if (sl->total) cudaFree(sl->total);
    return QTB_OK;
}

static int sliced_init_locked(qtb_ctx *ctx, qtb_sliced *sl, int nLanes) {
    sl->blobBytes = sl->lanes[0]->inBlobBytes;
    CU(cudaSetDevice(ctx->device));
    for (int l = 0; l < nLanes; l++) {
        cudaStream_t s = nullptr; cudaEvent_t e = nullptr;
        CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        sl->streams.push_back(s);
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sl->laneDone.push_back(e);
    }
    CU(cudaStreamCreateWithFlags(&sl->upStream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&sl->staged, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&sl->sumDone, cudaEventDisableTiming));
    CU(cudaMalloc((void **)&sl->laneAcc, sizeof(double2) * 8));
    CU(cudaMalloc((void **)&sl->total, sizeof(double2)));
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
    qtb_sliced *sl = new qtb_sliced_s();
    auto bail = [&](int st) {
        const std::string keep = g_lastError;
        for (qtb_plan *p : sl->lanes) qtb_plan_destroy(ctx, p);
        { std::lock_guard<std::mutex> lk(ctx->mu); sliced_destroy_locked(ctx, sl); }
        delete sl;
        g_lastError = keep;
        return st;
    };
    for (int l = 0; l < nLanes; l++) {
        qtb_plan *p = nullptr;
        const int st = qtb_plan_create_sliced(ctx, nInputs, inputRanks, nSteps, steps, nPrefix, &p);
        if (st != QTB_OK) return bail(st);
        if (p->outRank != 0) { sl->lanes.push_back(p); return bail(fail(QTB_ERR_INVALID, "a sliced plan must end in a scalar")); }
        sl->lanes.push_back(p);
    }
    int st = QTB_OK;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        st = sliced_init_locked(ctx, sl, nLanes);
    }
    if (st != QTB_OK) return bail(st);
    *out = sl;
    return QTB_OK;
}

int qtb_sliced_destroy(qtb_ctx *ctx, qtb_sliced *sl) {
    if (!sl) return QTB_OK;
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    { std::lock_guard<std::mutex> lk(ctx->mu); sliced_destroy_locked(ctx, sl); }
    for (qtb_plan *p : sl->lanes) qtb_plan_destroy(ctx, p);
    delete sl;
    return QTB_OK;
}

int qtb_sliced_stage(qtb_ctx *ctx, qtb_sliced *sl, int slot, const double *const *hostInputs) {
    if (!ctx || !sl || slot < 0 || slot > 65535 || !hostInputs) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    const qtb_plan *p0 = sl->lanes[0];
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
    const int R = (int)sl->lanes.size(), used = std::min(R, n);
    const bool useGraph = plan_graphs_enabled() && !ctx->trace;
    for (int l = 0; l < used; l++) {
        qtb_plan *pl = sl->lanes[l];
        const size_t nSegs = pl->segs.size(), pre = pl->prefixSegs;
        if (useGraph && !pl->partGraphsTried) {
            pl->partGraphsTried = true;
            if (pre >= 2) pl->graphPrefix = plan_capture(ctx, pl, 0, pre);
            if (nSegs - pre >= 2) pl->graphSuffix = plan_capture(ctx, pl, pre, nSegs);
        }
    }
    int preSteps = 0, preMicro = 0;
    {
        const qtb_plan *p0 = sl->lanes[0];
        for (size_t k = 0; k < p0->prefixSegs; k++) { preSteps += p0->segs[k].nSteps; if (p0->segs[k].micro) preMicro += p0->segs[k].nSteps; }
    }
    for (int l = 0; l < used; l++) {
        qtb_plan *pl = sl->lanes[l];
        cudaStream_t s = sl->streams[l];
        const size_t nSegs = pl->segs.size(), pre = pl->prefixSegs;
        if (sl->stagedValid) CU(cudaStreamWaitEvent(s, sl->staged, 0));
        bool first = true;
        for (int j = l; j < n; j += R) {
            CU(cudaMemcpyAsync(pl->inBlobDev, sl->slotDev[slots[j]], sl->blobBytes, cudaMemcpyDeviceToDevice, s));
            if (first && pre > 0) {
                if (useGraph && pl->graphPrefix) { CU(cudaGraphLaunch(pl->graphPrefix, s)); ctx->stats.launches += pl->launchesPrefix; }
                else ST(plan_enqueue(ctx, pl, s, 0, pre));
            }
            if (useGraph && pl->graphSuffix) { CU(cudaGraphLaunch(pl->graphSuffix, s)); ctx->stats.launches += pl->launches - pl->launchesPrefix; }
            else ST(plan_enqueue(ctx, pl, s, pre, nSegs));
            // the lane accumulator of the previous amplitude must have been folded into its total before it is overwritten
            if (first && sl->sumDoneValid) CU(cudaStreamWaitEvent(s, sl->sumDone, 0));
            k_acc_scalar<<<1, 32, 0, s>>>(sl->laneAcc + l, pl->outDev, first ? 1 : 0);
            ctx->stats.launches++;
            first = false;
        }
        CU(cudaGetLastError());
        CU(cudaEventRecord(sl->laneDone[l], s));
        CU(cudaStreamWaitEvent(ctx->stream, sl->laneDone[l], 0));
        ctx->stats.steps += preSteps; ctx->stats.micro_steps += preMicro; ctx->stats.units += pl->prefixUnits;
    }
    {
        const qtb_plan *p0 = sl->lanes[0];
        ctx->stats.steps += (long long)n * (p0->nSteps - preSteps);
        ctx->stats.micro_steps += (long long)n * (p0->nMicroSteps - preMicro);
        ctx->stats.units += (long long)n * (p0->units - p0->prefixUnits);
    }
    k_sum_scalars<<<1, 32, 0, ctx->stream>>>(sl->total, sl->laneAcc, used);       // used == 0: this rank owns no slice, total = 0
    CU(cudaGetLastError());
    ctx->stats.launches++;
    CU(cudaEventRecord(sl->sumDone, ctx->stream));
    sl->sumDoneValid = true;
    if (allreduce) {
        int r = g_nccl.AllReduce(sl->total, sl->total, 2, /*ncclDouble*/ 8, /*ncclSum*/ 0, ctx->comm, ctx->stream);
        if (r != 0) return fail(QTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    }
    qtb_scalar_read *rd = nullptr;
    ST(scalar_read_enqueue(ctx, sl->total, &rd));
    *out = rd;
    return QTB_OK;
}

int qtb_sliced_lanes(qtb_sliced *sl) { return sl ? (int)sl->lanes.size() : 0; }
long long qtb_sliced_units(qtb_sliced *sl) { return sl && !sl->lanes.empty() ? sl->lanes[0]->units : 0; }
long long qtb_sliced_prefix_units(qtb_sliced *sl) { return sl && !sl->lanes.empty() ? sl->lanes[0]->prefixUnits : 0; }
int qtb_sliced_launches(qtb_sliced *sl, int *prefixLaunches) {
    if (!sl || sl->lanes.empty()) return 0;
    if (prefixLaunches) *prefixLaunches = sl->lanes[0]->launchesPrefix;
    return sl->lanes[0]->launches;
}

}  // extern "C"
