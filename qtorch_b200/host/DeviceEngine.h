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
#include <cstdlib>
#include <limits>
#include <mutex>
#include <string>
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

}  // namespace device
}  // namespace qtorch
