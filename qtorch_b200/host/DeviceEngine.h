// qtorch_b200/host/DeviceEngine.h -- the only place the host headers touch libqtorch_b200.so.
//
// One process-wide engine context (device chosen by QTORCH_DEVICE, else LOCAL_RANK, else 0) shared by
// every Network: Node storage lives in its pool, Network::ContractIndices enqueues steps on its stream.
// Status codes of the C ABI are mapped back onto the reference's exception types
// (/root/reference/src/Exceptions.h:36-46) so that main.cpp's catch(std::exception&) blocks behave as before.
//
// "Plan-only" mode (QTORCH_PLAN_ONLY=1 or Engine::SetPlanOnly(true)) never touches the device: the host
// bookkeeping runs and the contraction plan is recorded, but no value is computed.  It exists so the
// planner logic can be tested and plans exported on machines without a GPU; it is NOT a CPU fallback --
// tensors have no data in that mode and every value read returns NaN.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <chrono>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <functional>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/qtorch_b200.h"
#include "Exceptions.h"

namespace qtorch {
namespace device {

inline void check(int status) {
    if (status == QTB_OK) return;
    const std::string detail = std::string(qtb_status_string(status)) + " (" + qtb_last_error() + ")";
    switch (status) {
        case QTB_ERR_INVALID:
        case QTB_ERR_EMPTY_INPUT: throw InvalidFunctionInput();
        case QTB_ERR_NO_DEVICE: throw DeviceUnavailable(detail);
        default: throw DeviceUnavailable(detail);
    }
}

class Engine {
public:
    static Engine &Get() {
        static Engine e;
        return e;
    }
    static bool &PlanOnlyFlag() {
        static bool flag = [] {
            const char *e = std::getenv("QTORCH_PLAN_ONLY");
            return e && std::atoi(e) != 0;
        }();
        return flag;
    }
    static void SetPlanOnly(bool on) { PlanOnlyFlag() = on; }
    static bool PlanOnly() { return PlanOnlyFlag(); }

    qtb_ctx *ctx() {
        std::lock_guard<std::mutex> lk(mMu);
        if (!mCtx) {
            int dev = 0;
            if (const char *e = std::getenv("QTORCH_DEVICE")) dev = std::atoi(e);
            else if (const char *e2 = std::getenv("LOCAL_RANK")) dev = std::atoi(e2);
            check(qtb_ctx_create(dev, &mCtx));       // throws DeviceUnavailable: there is no CPU path
        }
        return mCtx;
    }
    bool alive() const { return mCtx != nullptr; }
    void shutdown() {
        std::lock_guard<std::mutex> lk(mMu);
        if (mCtx) { qtb_ctx_destroy(mCtx); mCtx = nullptr; }
    }
    ~Engine() { /* leaked on purpose at process exit: CUDA may already be torn down */ }

private:
    Engine() {}
    std::mutex mMu;
    qtb_ctx *mCtx{nullptr};
};

// One process per GPU, launched by any launcher that exports RANK / WORLD_SIZE / LOCAL_RANK (torchrun --no-python,
// mpirun wrappers, a shell loop): joins the job's NCCL communicator and hands back the scalar allreduce the term
// dispatcher needs (edges dealt round-robin, ONE ncclAllReduce of the partial objective per evaluation).
// The NCCL unique id travels through a file: rank 0 writes it (atomically), the others wait for it.  Path:
// QTORCH_NCCL_ID_FILE, else /tmp/qtorch-<uid>/nccl_id.<MASTER_PORT or 0>.<QTORCH_JOB_ID or launcher pid>.  Single-process runs (WORLD_SIZE unset or 1) skip
// all of this and get an empty functor.
struct Job {
    int rank = 0, world = 1;
    std::function<void(double *, int)> allreduce;      // in-place sum of n doubles over all ranks; empty when world == 1

    static Job FromEnvironment() {
        Job job;
        if (const char *w = std::getenv("WORLD_SIZE")) job.world = std::max(1, std::atoi(w));
        if (const char *r = std::getenv("RANK")) job.rank = std::atoi(r);
        if (job.world == 1) { job.rank = 0; return job; }
        if (job.rank < 0 || job.rank >= job.world) throw DeviceUnavailable("RANK outside [0, WORLD_SIZE)");
        // The id file lives in a per-user directory (mode 0700) and carries a job nonce in its name -- QTORCH_JOB_ID, else
        // the launcher's pid (all ranks of one torchrun / shell loop share their parent) -- so neither another user nor a
        // job that died a moment ago can hand this job a stale id.  Rank 0 removes any leftover and creates the file
        // exclusively (O_EXCL) before publishing it by rename.
        std::string path;
        if (const char *f = std::getenv("QTORCH_NCCL_ID_FILE")) path = f;
        else {
            const std::string dir = "/tmp/qtorch-" + std::to_string(static_cast<long>(getuid()));
            mkdir(dir.c_str(), 0700);
            struct stat ds;
            if (stat(dir.c_str(), &ds) != 0 || ds.st_uid != getuid() || (ds.st_mode & 077) != 0)
                throw DeviceUnavailable("NCCL id directory " + dir + " is not private to this user");
            const char *port = std::getenv("MASTER_PORT"), *jid = std::getenv("QTORCH_JOB_ID");
            path = dir + "/nccl_id." + (port ? port : "0") + "." + (jid ? std::string(jid) : std::to_string(static_cast<long>(getppid())));
        }
        char id[QTB_UNIQUE_ID_BYTES];
        if (job.rank == 0) {
            check(qtb_comm_unique_id(id));
            const std::string tmp = path + ".tmp." + std::to_string(static_cast<long>(getpid()));
            std::remove(path.c_str());
            std::remove(tmp.c_str());
            const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0600);
            if (fd < 0 || write(fd, id, QTB_UNIQUE_ID_BYTES) != QTB_UNIQUE_ID_BYTES) { if (fd >= 0) close(fd); throw DeviceUnavailable("cannot write the NCCL id at " + tmp); }
            close(fd);
            if (std::rename(tmp.c_str(), path.c_str()) != 0) throw DeviceUnavailable("cannot publish the NCCL id at " + path);
        } else {
            // belt and braces: an id older than 30 s before this rank started is not from this launch
            const time_t notBefore = time(nullptr) - 30;
            bool got = false;
            for (int tries = 0; tries < 1200 && !got; ++tries) {            // up to 60 s
                struct stat st;
                std::ifstream in(path, std::ios::binary);
                if (stat(path.c_str(), &st) == 0 && st.st_uid == getuid() && st.st_mtime >= notBefore && in && in.read(id, QTB_UNIQUE_ID_BYTES)) got = true;
                else std::this_thread::sleep_for(std::chrono::milliseconds(50));
            }
            if (!got) throw DeviceUnavailable("no NCCL id from rank 0 at " + path);
        }
        qtb_ctx *ctx = Engine::Get().ctx();
        check(qtb_comm_init(ctx, job.world, job.rank, id));
        // everyone has joined once the first collective returns: rank 0 can take the id file away
        std::vector<double> probe(2, 1.0);
        check(qtb_allreduce_sum(ctx, probe.data(), 1));
        if (job.rank == 0) std::remove(path.c_str());
        job.allreduce = [ctx](double *v, int n) {
            std::vector<double> buf(2 * static_cast<size_t>(n), 0.0);        // the C ABI reduces complex scalars
            for (int i = 0; i < n; ++i) buf[2 * i] = v[i];
            check(qtb_allreduce_sum(ctx, buf.data(), n));
            for (int i = 0; i < n; ++i) v[i] = buf[2 * i];
        };
        return job;
    }
};

}  // namespace device
}  // namespace qtorch
