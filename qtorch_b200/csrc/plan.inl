// qtorch_b200/csrc/plan.inl -- compiled contraction plans (included at the end of engine.cu).
//
// A plan is the list of (A, B, shared legs) steps that a sequence of Network::ContractNodes calls
// produces (/root/reference/src/Network.h:715-864; the reference's own record of it is the list of
// mCreatedFrom pairs, Network.h:853-857).  Compiling it once assigns pooled device buffers with
// liveness-based reuse, pre-builds the grouped micro-step blobs and captures the launch sequence in
// a CUDA graph, so repeated evaluations (QAOA terms per COBYLA iteration, slices) cost a handful of
// launches each.

struct PlanSeg {
    bool fused = false;           // DMMA step + the inner product that follows, one kernel (g2 / tIsA / D describe the second step)
    StepGeom g2; bool tIsA = false; const double2 *D = nullptr;
    bool micro = false;
    uint32_t microIndex = 0;      // index into segOffsetsDev
    long long microUnits = 0;     // sum 4^(rC+k) over the segment's steps (heavy segments run on a cluster, engine.cu launch_micro_group)
    int nSteps = 0;
    StepGeom g; int kind = 0; GettChoice gc{0, false};
    const double2 *A = nullptr, *B = nullptr; double2 *C = nullptr;
};

struct qtb_plan_s {
    Pool pool;
    int nInputs = 0;
    std::vector<int> inputRanks;
    std::vector<double2 *> inputDev;
    std::vector<size_t> inputBlobOff;       // (size_t)-1 for big inputs living in the pool
    uint8_t *inBlobHost = nullptr, *inBlobDev = nullptr; size_t inBlobBytes = 0;
    cudaEvent_t inEvent = nullptr; bool inEventValid = false;
    uint8_t *microBlobDev = nullptr; uint64_t *segOffsetsDev = nullptr; size_t microBlobBytes = 0;
    std::vector<PlanSeg> segs;
    double2 *outDev = nullptr; int outRank = 0;
    long long units = 0; int nSteps = 0, nMicroSteps = 0; int launches = 0;
    cudaGraphExec_t graph = nullptr; bool graphTried = false;
    std::vector<uint8_t *> slotDev;          // extra resident copies of the small-input blob
    double2 *partialsDev = nullptr;          // private split-K / fused-dot partials (plans of one batch run on parallel streams)
    // slot-invariant prefix (qtb_plan_create_sliced): segs[0..prefixSegs) do not depend on the inputs that differ between
    // slots, so qtb_plan_run_slots runs them once and only segs[prefixSegs..) per slot; their results stay live.
    int nPrefixSteps = 0; size_t prefixSegs = 0; int launchesPrefix = 0; long long prefixUnits = 0;
    cudaGraphExec_t graphPrefix = nullptr, graphSuffix = nullptr; bool partGraphsTried = false;
    // result buffers in allocation order, split at the prefix / suffix boundary: another replica of the same plan can be
    // compiled onto the SAME addresses for either phase (the sliced executor shares prefix results between lanes and
    // suffix scratch between the two amplitudes a lane has in flight)
    std::vector<void *> prefixAllocs, suffixAllocs;
    bool prefixOnly = false;
};

static bool plan_graphs_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("QTB_PLAN_GRAPH"); v = (e && !atoi(e)) ? 0 : 1; }
    return v == 1;
}

static int plan_create_replica(qtb_ctx *ctx, int nInputs, const int *inputRanks, int nSteps, const qtb_plan_step *steps, int nPrefix,
                               const qtb_plan *prefixDonor, const qtb_plan *suffixDonor, bool prefixOnly, qtb_plan **out);

static int plan_enqueue(qtb_ctx *ctx, qtb_plan *pl, cudaStream_t s, size_t segBegin = 0, size_t segEnd = (size_t)-1) {
    if (segEnd > pl->segs.size()) segEnd = pl->segs.size();
    struct ScratchScope {                      // launches of this plan write their partial sums into the plan's own buffer
        qtb_ctx *c; double2 *saved;
        ScratchScope(qtb_ctx *ctx, double2 *p) : c(ctx), saved(ctx->scratchOverride) { if (p) c->scratchOverride = p; }
        ~ScratchScope() { c->scratchOverride = saved; }
    } scope(ctx, pl->partialsDev);
    for (size_t si = segBegin; si < segEnd; si++) {
        const PlanSeg &sg = pl->segs[si];
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (ctx->trace) { CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventRecord(e0, s)); }
        if (sg.micro) {
            ST(launch_micro_group(ctx, pl->microBlobDev, pl->segOffsetsDev + sg.microIndex, sg.microUnits, s));
            ctx->stats.launches++;
        } else if (sg.fused) {
            ST(enqueue_fused(ctx, sg.g, sg.gc, sg.A, sg.B, sg.g2, sg.tIsA, sg.D, sg.C, s));
        } else {
            ST(enqueue_big(ctx, sg.g, sg.kind, sg.gc, sg.A, sg.B, sg.C, s));
        }
        if (ctx->trace) {
            CU(cudaEventRecord(e1, s));
            ctx->traceRecs.push_back({e0, e1, sg.micro ? 0 : sg.g.rA, sg.micro ? 0 : sg.g.rB, sg.micro ? sg.nSteps : sg.g.k,
                                      sg.micro ? KIND_MICRO : sg.fused ? KIND_FUSED : sg.kind});
        }
    }
    return QTB_OK;
}

extern "C" {

int qtb_plan_create(qtb_ctx *ctx, int nInputs, const int *inputRanks, int nSteps, const qtb_plan_step *steps, qtb_plan **out) {
    return qtb_plan_create_sliced(ctx, nInputs, inputRanks, nSteps, steps, 0, out);
}

int qtb_plan_create_sliced(qtb_ctx *ctx, int nInputs, const int *inputRanks, int nSteps, const qtb_plan_step *steps, int nPrefix, qtb_plan **out) {
    return plan_create_replica(ctx, nInputs, inputRanks, nSteps, steps, nPrefix, nullptr, nullptr, false, out);
}

}  // extern "C"

// prefixDonor / suffixDonor: replicas of the same plan whose result buffers this one re-uses for that phase (same
// allocation order => same tensor-to-buffer map).  prefixOnly: compile the invariant prefix alone.
static int plan_create_replica(qtb_ctx *ctx, int nInputs, const int *inputRanks, int nSteps, const qtb_plan_step *steps, int nPrefix,
                               const qtb_plan *prefixDonor, const qtb_plan *suffixDonor, bool prefixOnly, qtb_plan **out) {
    if (!ctx || !out || nInputs < 0 || nSteps < 1 || (nInputs > 0 && !inputRanks) || !steps) return fail(QTB_ERR_INVALID, "bad plan arguments");
    if (nPrefix < 0 || nPrefix > nSteps) return fail(QTB_ERR_INVALID, "bad invariant-prefix length");
    if (nPrefix == nSteps) nPrefix = 0;              // nothing varies: an ordinary plan
    *out = nullptr;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    const int nT = nInputs + nSteps;
    std::vector<int> rank(nT, -1);
    std::vector<char> consumed(nT, 0);
    for (int i = 0; i < nInputs; i++) {
        if (inputRanks[i] < 0 || inputRanks[i] > QTB_MAX_RANK) return fail(QTB_ERR_INVALID, "input rank out of range");
        rank[i] = inputRanks[i];
    }
    // ---- validate + geometry
    std::vector<StepGeom> geoms(nSteps);
    for (int i = 0; i < nSteps; i++) {
        const qtb_plan_step &s = steps[i];
        if (s.a < 0 || s.b < 0 || s.a >= nInputs + i || s.b >= nInputs + i || s.a == s.b) return fail(QTB_ERR_INVALID, "plan step references an unknown tensor");
        if (consumed[s.a] || consumed[s.b]) return fail(QTB_ERR_EMPTY_INPUT, "plan step re-uses an already contracted tensor (Network.h:719-723)");
        int pa[QTB_MAXR], pb[QTB_MAXR];
        if (s.k < 0 || s.k > QTB_MAXR) return fail(QTB_ERR_INVALID, "bad shared-leg count");
        for (int j = 0; j < s.k; j++) { pa[j] = s.pos_a[j]; pb[j] = s.pos_b[j]; }
        ST(make_geom(rank[s.a], rank[s.b], s.k, pa, pb, geoms[i]));
        rank[nInputs + i] = geoms[i].rC;
        consumed[s.a] = consumed[s.b] = 1;
    }
    qtb_plan *pl = new qtb_plan_s();
    auto bail = [&](int st) { pl->pool.destroy(); if (pl->inBlobHost) cudaFreeHost(pl->inBlobHost); if (pl->inBlobDev) cudaFree(pl->inBlobDev);
                              if (pl->partialsDev) cudaFree(pl->partialsDev);
                              if (pl->microBlobDev) cudaFree(pl->microBlobDev); if (pl->segOffsetsDev) cudaFree(pl->segOffsetsDev); delete pl; return st; };
    pl->nInputs = nInputs; pl->inputRanks.assign(inputRanks, inputRanks + nInputs);
    pl->inputDev.assign(nInputs, nullptr); pl->inputBlobOff.assign(nInputs, (size_t)-1);
    // ---- inputs: small ones share one staging blob (one H2D per evaluation)
    size_t blobBytes = 0;
    for (int i = 0; i < nInputs; i++) {
        if (rank[i] <= 5) { pl->inputBlobOff[i] = blobBytes; blobBytes += std::max<size_t>(Pool::bytes(rank[i]), 256); }
    }
    pl->inBlobBytes = blobBytes;
    if (blobBytes) {
        if (cudaMallocHost((void **)&pl->inBlobHost, blobBytes) != cudaSuccess || cudaMalloc((void **)&pl->inBlobDev, blobBytes) != cudaSuccess)
            return bail(fail(QTB_ERR_OOM, "plan input staging allocation failed"));
    }
    std::vector<double2 *> dev(nT, nullptr);
    for (int i = 0; i < nInputs; i++) {
        if (pl->inputBlobOff[i] != (size_t)-1) dev[i] = (double2 *)(pl->inBlobDev + pl->inputBlobOff[i]);
        else { void *p = nullptr; int st = pl->pool.alloc(rank[i], &p); if (st != QTB_OK) return bail(st); dev[i] = (double2 *)p; }
        pl->inputDev[i] = dev[i];
    }
    // ---- walk the steps: buffers, kernel families, micro segments
    std::vector<std::vector<uint8_t>> microBlobs;
    std::vector<PendingStep> cur;
    std::vector<uint32_t> levelOf(nT, 0);           // level (+1) of the pending micro-step that produces a tensor, 0 = ready
    std::vector<std::pair<int, void *>> deferred;
    auto closeMicro = [&]() {
        if (cur.empty()) return;
        PlanSeg sg; sg.micro = true; sg.microIndex = (uint32_t)microBlobs.size(); sg.nSteps = (int)cur.size();
        for (const auto &ps : cur) sg.microUnits += 1ll << (2 * (ps.st.rC + ps.st.k));
        pl->segs.push_back(sg);
        microBlobs.emplace_back();                   // assembled once all device addresses are known (below)
        // keep the raw steps: stash them in the blob vector as bytes for now
        microBlobs.back().resize(cur.size() * sizeof(PendingStep));
        memcpy(microBlobs.back().data(), cur.data(), cur.size() * sizeof(PendingStep));
        cur.clear();
        std::fill(levelOf.begin(), levelOf.end(), 0);
        for (auto &f : deferred) pl->pool.release(f.first, f.second);
        deferred.clear();
    };
    pl->nPrefixSteps = nPrefix;
    pl->prefixOnly = prefixOnly;
    if (prefixOnly && nPrefix < 1) return bail(fail(QTB_ERR_INVALID, "a prefix-only replica needs an invariant prefix"));
    // a tensor made by the invariant prefix (or a plan input) must survive every slot's pass over the suffix
    auto releasable = [&](int t, int atStep) { return t >= nInputs && !(atStep >= nPrefix && t < nInputs + nPrefix); };
    // result buffers: from this plan's pool, or -- phase by phase -- the addresses a donor replica got for the same request
    std::vector<char> borrowed(nT, 0);                   // tensor lives in a donor's buffer: never released into OUR pool
    size_t replayPre = 0, replaySuf = 0;
    auto allocResult = [&](int step, int rk, int tensorId, void **outp) -> int {
        const bool inPrefix = step < nPrefix;
        const qtb_plan *donor = inPrefix ? prefixDonor : suffixDonor;
        if (donor) {
            const std::vector<void *> &rec = inPrefix ? donor->prefixAllocs : donor->suffixAllocs;
            size_t &cur = inPrefix ? replayPre : replaySuf;
            if (cur >= rec.size()) return fail(QTB_ERR_INVALID, "donor replica does not match this plan");
            *outp = rec[cur++];
            if (tensorId >= 0) borrowed[tensorId] = 1;
        } else {
            ST(pl->pool.alloc(rk, outp));
        }
        (inPrefix ? pl->prefixAllocs : pl->suffixAllocs).push_back(*outp);
        return QTB_OK;
    };
    auto releaseTensor = [&](int t, void *ptr) { if (!borrowed[t]) pl->pool.release(rank[t], ptr); };
    const int nCompiled = prefixOnly ? nPrefix : nSteps;
    for (int i = 0; i < nCompiled; i++) {
        const qtb_plan_step &s = steps[i];
        const StepGeom &g = geoms[i];
        GettChoice gc{0, false};
        const int kind = choose_kind(g, gc, ctx->microLog4);
        pl->units += (long long)g.units();
        if (nPrefix > 0 && i == nPrefix) {
            closeMicro(); pl->prefixSegs = pl->segs.size(); pl->prefixUnits = pl->units - (long long)g.units();
            // suffix results never land in memory the prefix phase has used: the prefix of the NEXT amplitude may already be
            // running (on its own copy of the prefix buffers) while this amplitude's slices are still being contracted
            for (auto &fl : pl->pool.freeList) fl.clear();
        }
        // fusion: this DMMA step followed by the inner product of its result with another tensor
        if (i + 1 < nSteps && i + 1 != nPrefix && (steps[i + 1].a == nInputs + i || steps[i + 1].b == nInputs + i)) {
            const bool tIsA = steps[i + 1].a == nInputs + i;
            GettChoice gc2{0, false};
            if (choose_kind(geoms[i + 1], gc2, ctx->microLog4) == KIND_REDUCE && fusable_pair(g, kind, gc, geoms[i + 1], tIsA)) {
                closeMicro();
                void *outp = nullptr;
                { int st = allocResult(i + 1, 0, nInputs + i + 1, &outp); if (st != QTB_OK) return bail(st); }
                const int dId = tIsA ? steps[i + 1].b : steps[i + 1].a;
                PlanSeg sg; sg.micro = false; sg.fused = true; sg.g = g; sg.kind = KIND_GETT; sg.gc = gc; sg.nSteps = 2;
                sg.A = dev[s.a]; sg.B = dev[s.b]; sg.C = (double2 *)outp; sg.g2 = geoms[i + 1]; sg.tIsA = tIsA; sg.D = dev[dId];
                pl->segs.push_back(sg);
                dev[nInputs + i] = nullptr;                       // the intermediate is never materialised
                dev[nInputs + i + 1] = (double2 *)outp;
                pl->units += (long long)geoms[i + 1].units();
                if (releasable(s.a, i)) releaseTensor(s.a, dev[s.a]);
                if (releasable(s.b, i)) releaseTensor(s.b, dev[s.b]);
                if (releasable(dId, i)) releaseTensor(dId, dev[dId]);
                ++i;                                              // the inner-product step is consumed
                continue;
            }
        }
        void *cp = nullptr;
        { int st = allocResult(i, g.rC, nInputs + i, &cp); if (st != QTB_OK) return bail(st); }
        dev[nInputs + i] = (double2 *)cp;
        if (kind == KIND_MICRO) {
            PendingStep ps;
            make_devstep(g, dev[s.a], dev[s.b], dev[nInputs + i], KIND_MICRO, ps.st);
            ps.level = std::max(levelOf[s.a], levelOf[s.b]);
            cur.push_back(ps);
            levelOf[nInputs + i] = ps.level + 1;
            pl->nMicroSteps++;
            if (releasable(s.a, i) && !borrowed[s.a]) deferred.push_back({rank[s.a], (void *)dev[s.a]});
            if (releasable(s.b, i) && !borrowed[s.b]) deferred.push_back({rank[s.b], (void *)dev[s.b]});
        } else {
            closeMicro();
            PlanSeg sg; sg.micro = false; sg.g = g; sg.kind = kind; sg.gc = gc; sg.nSteps = 1;
            sg.A = dev[s.a]; sg.B = dev[s.b]; sg.C = dev[nInputs + i];
            pl->segs.push_back(sg);
            if (releasable(s.a, i)) releaseTensor(s.a, dev[s.a]);
            if (releasable(s.b, i)) releaseTensor(s.b, dev[s.b]);
        }
    }
    closeMicro();
    if (prefixOnly) { pl->prefixSegs = pl->segs.size(); pl->prefixUnits = pl->units; }
    pl->nSteps = nCompiled;
    pl->outDev = prefixOnly ? nullptr : dev[nT - 1]; pl->outRank = rank[nT - 1];
    pl->launches = 0;
    for (size_t si = 0; si < pl->segs.size(); si++) {
        const PlanSeg &sg = pl->segs[si];
        const int l = (!sg.micro && (sg.kind == KIND_REDUCE || sg.fused)) ? 2 : 1;
        pl->launches += l;
        if (si < pl->prefixSegs) pl->launchesPrefix += l;
    }
    // ---- assemble micro blobs into one device allocation
    if (!microBlobs.empty()) {
        std::vector<uint64_t> offs(microBlobs.size());
        std::vector<std::vector<uint8_t>> built(microBlobs.size());
        size_t total = 0;
        for (size_t m = 0; m < microBlobs.size(); m++) {
            std::vector<PendingStep> st(microBlobs[m].size() / sizeof(PendingStep));
            memcpy(st.data(), microBlobs[m].data(), microBlobs[m].size());
            build_micro_blob(st, {}, {}, built[m], nullptr, m == 0 ? pl->inBlobDev : nullptr, m == 0 ? pl->inBlobBytes : 0);
            offs[m] = total;
            total += (built[m].size() + 255) & ~(size_t)255;
        }
        pl->microBlobBytes = total;
        if (cudaMalloc((void **)&pl->microBlobDev, total) != cudaSuccess || cudaMalloc((void **)&pl->segOffsetsDev, offs.size() * 8) != cudaSuccess)
            return bail(fail(QTB_ERR_OOM, "plan blob allocation failed"));
        std::vector<uint8_t> hostAll(total, 0);
        for (size_t m = 0; m < built.size(); m++) memcpy(hostAll.data() + offs[m], built[m].data(), built[m].size());
        if (cudaMemcpy(pl->microBlobDev, hostAll.data(), total, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(pl->segOffsetsDev, offs.data(), offs.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(QTB_ERR_CUDA, "plan blob upload failed"));
    }
    if (cudaEventCreateWithFlags(&pl->inEvent, cudaEventDisableTiming) != cudaSuccess) return bail(fail(QTB_ERR_CUDA, "event"));
    for (const PlanSeg &sg : pl->segs)
        if (!sg.micro && (sg.kind == KIND_REDUCE || sg.fused) && !pl->partialsDev &&
            cudaMalloc((void **)&pl->partialsDev, (size_t)REDUCE_MAX_BLOCKS * 16 * sizeof(double2)) != cudaSuccess)
            return bail(fail(QTB_ERR_OOM, "plan partial-sum buffer allocation failed"));
    *out = pl;
    return QTB_OK;
}

extern "C" {

int qtb_plan_destroy(qtb_ctx *ctx, qtb_plan *pl) {
    if (!pl) return QTB_OK;
    if (!ctx) return fail(QTB_ERR_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (pl->graph) cudaGraphExecDestroy(pl->graph);
    if (pl->graphPrefix) cudaGraphExecDestroy(pl->graphPrefix);
    if (pl->graphSuffix) cudaGraphExecDestroy(pl->graphSuffix);
    pl->pool.destroy();
    if (pl->inBlobHost) cudaFreeHost(pl->inBlobHost);
    if (pl->inBlobDev) cudaFree(pl->inBlobDev);
    if (pl->microBlobDev) cudaFree(pl->microBlobDev);
    if (pl->segOffsetsDev) cudaFree(pl->segOffsetsDev);
    if (pl->inEvent) cudaEventDestroy(pl->inEvent);
    if (pl->partialsDev) cudaFree(pl->partialsDev);
    for (uint8_t *p : pl->slotDev) if (p) cudaFree(p);
    delete pl;
    return QTB_OK;
}

static int plan_upload_locked(qtb_ctx *ctx, qtb_plan *pl, const double *const *hostInputs) {
    if (pl->nInputs > 0 && !hostInputs) return fail(QTB_ERR_INVALID, "null inputs");
    ST(ensure_device(ctx));
    if (pl->inEventValid) CU(cudaEventSynchronize(pl->inEvent));     // previous H2D finished reading the pinned blob
    for (int i = 0; i < pl->nInputs; i++) {
        if (!hostInputs[i]) return fail(QTB_ERR_EMPTY_INPUT, "null input tensor");
        const size_t b = Pool::bytes(pl->inputRanks[i]);
        if (pl->inputBlobOff[i] != (size_t)-1) memcpy(pl->inBlobHost + pl->inputBlobOff[i], hostInputs[i], b);
        else CU(cudaMemcpyAsync(pl->inputDev[i], hostInputs[i], b, cudaMemcpyHostToDevice, ctx->stream));
        ctx->stats.bytes_h2d += (long long)b;
    }
    if (pl->inBlobBytes) {
        CU(cudaMemcpyAsync(pl->inBlobDev, pl->inBlobHost, pl->inBlobBytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaEventRecord(pl->inEvent, ctx->stream));
        pl->inEventValid = true;
    }
    return QTB_OK;
}

// capture segs[segBegin, segEnd) into an executable graph; nullptr (and no error) when capture is not possible
static cudaGraphExec_t plan_capture(qtb_ctx *ctx, qtb_plan *pl, size_t segBegin, size_t segEnd) {
    cudaGraph_t g = nullptr;
    cudaGraphExec_t exec = nullptr;
    const long long launchesBefore = ctx->stats.launches;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        int st = plan_enqueue(ctx, pl, ctx->stream, segBegin, segEnd);
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        ctx->stats.launches = launchesBefore;
        if (st == QTB_OK && e == cudaSuccess && g) {
            if (cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) { exec = nullptr; cudaGetLastError(); }
        } else cudaGetLastError();
        if (g) cudaGraphDestroy(g);
    } else cudaGetLastError();
    return exec;
}

static int plan_run_locked(qtb_ctx *ctx, qtb_plan *pl, cudaStream_t stream = nullptr) {
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    if (!stream) stream = ctx->stream;
    const bool useGraph = plan_graphs_enabled() && !ctx->trace && pl->segs.size() >= 3;
    if (useGraph && !pl->graphTried) {
        pl->graphTried = true;
        pl->graph = plan_capture(ctx, pl, 0, pl->segs.size());
    }
    if (useGraph && pl->graph) {
        CU(cudaGraphLaunch(pl->graph, stream));
        ctx->stats.launches += pl->launches;
    } else {
        ST(plan_enqueue(ctx, pl, stream));
    }
    ctx->stats.steps += pl->nSteps;
    ctx->stats.micro_steps += pl->nMicroSteps;
    ctx->stats.units += pl->units;
    return QTB_OK;
}

static int plan_read_locked(qtb_ctx *ctx, qtb_plan *pl, double *hostOut) {
    const size_t b = Pool::bytes(pl->outRank);
    if (pl->outRank == 0) {
        CU(cudaMemcpyAsync(ctx->scalarPinned, pl->outDev, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        hostOut[0] = ctx->scalarPinned[0]; hostOut[1] = ctx->scalarPinned[1];
    } else {
        CU(cudaMemcpyAsync(hostOut, pl->outDev, b, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->stats.bytes_d2h += (long long)b;
    return QTB_OK;
}

int qtb_plan_upload_inputs(qtb_ctx *ctx, qtb_plan *pl, const double *const *hostInputs) {
    if (!ctx || !pl) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return plan_upload_locked(ctx, pl, hostInputs);
}
int qtb_plan_run_device(qtb_ctx *ctx, qtb_plan *pl) {
    if (!ctx || !pl) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return plan_run_locked(ctx, pl);
}
int qtb_plan_read_output(qtb_ctx *ctx, qtb_plan *pl, double *hostOut) {
    if (!ctx || !pl || !hostOut) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    return plan_read_locked(ctx, pl, hostOut);
}
int qtb_plan_run_host(qtb_ctx *ctx, qtb_plan *pl, const double *const *hostInputs, double *hostOut) {
    if (!ctx || !pl || !hostOut) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(plan_upload_locked(ctx, pl, hostInputs));
    ST(plan_run_locked(ctx, pl));
    return plan_read_locked(ctx, pl, hostOut);
}
int qtb_plan_stage_inputs(qtb_ctx *ctx, qtb_plan *pl, int slot, const double *const *hostInputs) {
    if (!ctx || !pl || slot < 0 || slot > 4096 || !hostInputs) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    for (int i = 0; i < pl->nInputs; i++)
        if (pl->inputBlobOff[i] == (size_t)-1) return fail(QTB_ERR_UNSUPPORTED, "input slots need every input to have rank <= 5");
    if ((size_t)slot >= pl->slotDev.size()) pl->slotDev.resize(slot + 1, nullptr);
    if (!pl->slotDev[slot]) CU(cudaMalloc((void **)&pl->slotDev[slot], std::max<size_t>(pl->inBlobBytes, 256)));
    if (pl->inEventValid) CU(cudaEventSynchronize(pl->inEvent));
    for (int i = 0; i < pl->nInputs; i++) {
        if (!hostInputs[i]) return fail(QTB_ERR_EMPTY_INPUT, "null input tensor");
        memcpy(pl->inBlobHost + pl->inputBlobOff[i], hostInputs[i], Pool::bytes(pl->inputRanks[i]));
        ctx->stats.bytes_h2d += (long long)Pool::bytes(pl->inputRanks[i]);
    }
    CU(cudaMemcpyAsync(pl->slotDev[slot], pl->inBlobHost, pl->inBlobBytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaEventRecord(pl->inEvent, ctx->stream));
    pl->inEventValid = true;
    return QTB_OK;
}
int qtb_plan_run_device_slot(qtb_ctx *ctx, qtb_plan *pl, int slot) {
    if (!ctx || !pl) return fail(QTB_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (slot < 0 || (size_t)slot >= pl->slotDev.size() || !pl->slotDev[slot]) return fail(QTB_ERR_INVALID, "unknown input slot");
    ST(ensure_device(ctx));
    CU(cudaMemcpyAsync(pl->inBlobDev, pl->slotDev[slot], pl->inBlobBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return plan_run_locked(ctx, pl);
}
int qtb_plan_run_slots(qtb_ctx *ctx, qtb_plan *pl, const int *slots, int n, double *hostSum, double *hostEach) {
    if (!ctx || !pl || !slots || n < 1 || !hostSum) return fail(QTB_ERR_INVALID, "bad argument");
    if (pl->outRank != 0) return fail(QTB_ERR_INVALID, "slot sums need a scalar plan output");
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (int j = 0; j < n; j++)
        if (slots[j] < 0 || (size_t)slots[j] >= pl->slotDev.size() || !pl->slotDev[slots[j]]) return fail(QTB_ERR_INVALID, "unknown input slot");
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    const size_t nSegs = pl->segs.size(), pre = pl->prefixSegs;
    const bool useGraph = plan_graphs_enabled() && !ctx->trace;
    if (useGraph && !pl->partGraphsTried) {
        pl->partGraphsTried = true;
        if (pre >= 2) pl->graphPrefix = plan_capture(ctx, pl, 0, pre);
        if (nSegs - pre >= 2) pl->graphSuffix = plan_capture(ctx, pl, pre, nSegs);
    }
    if ((size_t)n > ctx->batchOutCap) {
        if (ctx->batchOut) cudaFreeHost(ctx->batchOut);
        ctx->batchOutCap = std::max<size_t>(256, (size_t)n * 2);
        CU(cudaMallocHost((void **)&ctx->batchOut, ctx->batchOutCap * 16));
    }
    for (int j = 0; j < n; j++) {
        CU(cudaMemcpyAsync(pl->inBlobDev, pl->slotDev[slots[j]], pl->inBlobBytes, cudaMemcpyDeviceToDevice, ctx->stream));
        if (j == 0 && pre > 0) {
            // slot-invariant steps: once per call, from the first slot's copy of the (shared) inputs
            if (useGraph && pl->graphPrefix) { CU(cudaGraphLaunch(pl->graphPrefix, ctx->stream)); ctx->stats.launches += pl->launchesPrefix; }
            else ST(plan_enqueue(ctx, pl, ctx->stream, 0, pre));
        }
        if (useGraph && pl->graphSuffix) { CU(cudaGraphLaunch(pl->graphSuffix, ctx->stream)); ctx->stats.launches += pl->launches - pl->launchesPrefix; }
        else ST(plan_enqueue(ctx, pl, ctx->stream, pre, nSegs));
        CU(cudaMemcpyAsync(ctx->batchOut + 2 * j, pl->outDev, 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    double sr = 0.0, si = 0.0;
    for (int j = 0; j < n; j++) {
        sr += ctx->batchOut[2 * j]; si += ctx->batchOut[2 * j + 1];
        if (hostEach) { hostEach[2 * j] = ctx->batchOut[2 * j]; hostEach[2 * j + 1] = ctx->batchOut[2 * j + 1]; }
    }
    hostSum[0] = sr; hostSum[1] = si;
    // work actually done: the prefix once, the suffix n times
    int preSteps = 0, preMicro = 0;
    for (size_t k = 0; k < pre; k++) {
        preSteps += pl->segs[k].nSteps;
        if (pl->segs[k].micro) preMicro += pl->segs[k].nSteps;
    }
    ctx->stats.steps += preSteps + (long long)n * (pl->nSteps - preSteps);
    ctx->stats.micro_steps += preMicro + (long long)n * (pl->nMicroSteps - preMicro);
    ctx->stats.units += pl->prefixUnits + (long long)n * (pl->units - pl->prefixUnits);
    ctx->stats.bytes_d2h += (long long)n * 16;
    return QTB_OK;
}
long long qtb_plan_prefix_units(qtb_plan *pl) { return pl ? pl->prefixUnits : 0; }
int qtb_plans_run_batched(qtb_ctx *ctx, qtb_plan *const *plans, int n, const double *const *const *hostInputs, double *hostOut) {
    if (!ctx || !plans || n < 1 || !hostInputs || !hostOut) return fail(QTB_ERR_INVALID, "bad argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    ST(ensure_device(ctx));
    ST(flush_locked(ctx));
    bool allMicro = true;
    for (int i = 0; i < n; i++) {
        if (!plans[i] || plans[i]->outRank != 0) return fail(QTB_ERR_INVALID, "batched plans must have scalar outputs");
        if (plans[i]->segs.size() != 1 || !plans[i]->segs[0].micro) allMicro = false;
        // one plan = one set of buffers and one graph: the same plan twice in a batch would race with itself
        for (int j = 0; j < i; j++) if (plans[j] == plans[i]) return fail(QTB_ERR_INVALID, "the same plan appears twice in one batch");
    }
    for (int i = 0; i < n; i++) ST(plan_upload_locked(ctx, plans[i], hostInputs[i]));
    if (allMicro) {
        // one CTA per plan: absolute blob addresses through the staging ring
        size_t off = 0;
        ST(ring_reserve(ctx, (size_t)n * 8, off));
        uint64_t *addr = reinterpret_cast<uint64_t *>(ctx->ringHost + off);
        for (int i = 0; i < n; i++) addr[i] = reinterpret_cast<uint64_t>(plans[i]->microBlobDev);
        CU(cudaMemcpyAsync(ctx->ringDev + off, ctx->ringHost + off, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        long long unitsTotal = 0;
        for (int i = 0; i < n; i++) unitsTotal += plans[i]->units;
        ST(launch_micro_plans(ctx, n, unitsTotal, reinterpret_cast<const uint64_t *>(ctx->ringDev + off), ctx->stream));
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ringEvent, ctx->stream));
        ctx->ringEventValid = true;
        ctx->stats.launches++;
        for (int i = 0; i < n; i++) { ctx->stats.steps += plans[i]->nSteps; ctx->stats.micro_steps += plans[i]->nMicroSteps; ctx->stats.units += plans[i]->units; }
    } else if (ctx->trace || n == 1) {
        for (int i = 0; i < n; i++) ST(plan_run_locked(ctx, plans[i]));
    } else {
        // independent plans with their own buffers: fork over side streams, join before the read-back.  Most of a small
        // plan is one single-CTA grouped launch, so plans side by side fill the SMs that one plan leaves idle.
        const int nAux = std::min(n, 16);
        while ((int)ctx->auxStreams.size() < nAux) {
            cudaStream_t a = nullptr; cudaEvent_t e = nullptr;
            CU(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
            ctx->auxStreams.push_back(a);
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->auxEvents.push_back(e);
        }
        if (!ctx->forkEvent) CU(cudaEventCreateWithFlags(&ctx->forkEvent, cudaEventDisableTiming));
        // graphs are captured on the main stream (first run only), before the fork
        for (int i = 0; i < n; i++)
            if (plan_graphs_enabled() && plans[i]->segs.size() >= 3 && !plans[i]->graphTried) {
                plans[i]->graphTried = true;
                plans[i]->graph = plan_capture(ctx, plans[i], 0, plans[i]->segs.size());
            }
        CU(cudaEventRecord(ctx->forkEvent, ctx->stream));
        for (int a = 0; a < nAux; a++) CU(cudaStreamWaitEvent(ctx->auxStreams[a], ctx->forkEvent, 0));
        int st = QTB_OK;
        for (int i = 0; i < n && st == QTB_OK; i++) st = plan_run_locked(ctx, plans[i], ctx->auxStreams[i % nAux]);
        for (int a = 0; a < nAux; a++) {              // always join: whatever was forked must be ordered before later main-stream work
            CU(cudaEventRecord(ctx->auxEvents[a], ctx->auxStreams[a]));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->auxEvents[a], 0));
        }
        ST(st);
    }
    // gather the n scalars: n tiny async copies into pinned memory, one wait
    if ((size_t)n > ctx->batchOutCap) {
        if (ctx->batchOut) cudaFreeHost(ctx->batchOut);
        ctx->batchOutCap = std::max<size_t>(256, (size_t)n * 2);
        CU(cudaMallocHost((void **)&ctx->batchOut, ctx->batchOutCap * 16));
    }
    for (int i = 0; i < n; i++) CU(cudaMemcpyAsync(ctx->batchOut + 2 * i, plans[i]->outDev, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    memcpy(hostOut, ctx->batchOut, (size_t)n * 16);
    ctx->stats.bytes_d2h += (long long)n * 16;
    return QTB_OK;
}
int qtb_plan_output_rank(qtb_plan *pl) { return pl ? pl->outRank : -1; }
long long qtb_plan_units(qtb_plan *pl) { return pl ? pl->units : 0; }
int qtb_plan_launches(qtb_plan *pl) { return pl ? pl->launches : 0; }

}  // extern "C"
