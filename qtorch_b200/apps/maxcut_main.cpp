// qtorch_b200/apps/maxcut_main.cpp -- `maxcutQAOA <graph.dgf> <p> 0 <angle file> [max evaluations]` on the B200 engine.
// Counterpart of /root/reference/src/maxcut.cpp:227-328, mode 0 (find the QAOA angles that maximise F_p).
//  * The objective is the reference's F_p (maxcut.cpp:162-204) evaluated by QaoaObjective: per-edge light-cone
//    networks planned once, all edges of this rank in one grouped launch per evaluation.
//  * The reference maximises with NLopt's COBYLA, no stopping criterion set (maxcut.cpp:211-213).  When the build finds the
//    NLopt tree the reference vendors (third-party; qtorch_b200/build.py compiles it out of tree into bin/_nlopt), this
//    front-end does exactly the same -- nlopt::opt(LN_COBYLA, 2p), set_max_objective, optimize from beta = 0.392699,
//    gamma = 0.785399 (maxcut.cpp:155-157) -- so the sequence of evaluated angles is the reference's (the objective values
//    agree to ~1e-15) and the angle file holds the last evaluated angles, as the reference leaves it (maxcut.cpp:199-202).
//    The reference is stopped by NLopt's round-off detection, which only its numerical noise triggers; here xtol_abs = 1e-9 ends
//    the run instead (plus an evaluation cap, optional argument, default 20000).  Without NLopt (QTB_HAVE_NLOPT undefined) a small Nelder-Mead ascent
//    with an evaluation cap takes its place and the binary says so.
//  * Multi-GPU: started once per GPU with RANK / WORLD_SIZE / LOCAL_RANK in the environment (e.g.
//    `torchrun --no-python --nproc-per-node N maxcutQAOA ...`), every process owns the edges e with e % WORLD_SIZE == RANK
//    and the partial objectives meet in one NCCL allreduce per evaluation (device::Job::FromEnvironment).  All ranks
//    run the same deterministic optimiser on the same reduced values; rank 0 alone writes files and reports.
//  * Modes 1 and 2 add the final bit-string sampler (maxcut.cpp:29-140): the n-qubit circuit is planned once with the
//    in-process min-fill ordering and the n conditional probabilities re-use one compiled device plan.
#include <sys/stat.h>
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../host/qtorch.hpp"
#include "../host/maxcut.h"
#ifdef QTB_HAVE_NLOPT
#include <nlopt.hpp>
#endif

static std::vector<double> nelderMeadMaximise(const std::function<double(const std::vector<double> &)> &f, std::vector<double> x0,
                                              int maxEvals, double step, int &evals) {
    const size_t n = x0.size();
    std::vector<std::vector<double>> simplex(n + 1, x0);
    for (size_t i = 0; i < n; ++i) simplex[i + 1][i] += step;
    std::vector<double> val(n + 1);
    evals = 0;
    for (size_t i = 0; i <= n; ++i) { val[i] = f(simplex[i]); ++evals; }
    auto combine = [&](const std::vector<double> &c, const std::vector<double> &w, double t) {
        std::vector<double> r(n);
        for (size_t i = 0; i < n; ++i) r[i] = c[i] + t * (c[i] - w[i]);
        return r;
    };
    while (evals < maxEvals) {
        std::vector<size_t> idx(n + 1);
        for (size_t i = 0; i <= n; ++i) idx[i] = i;
        std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return val[a] > val[b]; });      // best first (maximise)
        const size_t best = idx[0], worst = idx[n], second = idx[n - 1];
        if (std::abs(val[best] - val[worst]) < 1e-12) break;
        std::vector<double> centroid(n, 0.0);
        for (size_t i = 0; i < n; ++i) for (size_t d = 0; d < n; ++d) centroid[d] += simplex[idx[i]][d] / n;
        std::vector<double> refl = combine(centroid, simplex[worst], 1.0);
        const double fr = f(refl); ++evals;
        if (fr > val[best]) {
            std::vector<double> exp = combine(centroid, simplex[worst], 2.0);
            const double fe = f(exp); ++evals;
            if (fe > fr) { simplex[worst] = exp; val[worst] = fe; } else { simplex[worst] = refl; val[worst] = fr; }
        } else if (fr > val[second]) {
            simplex[worst] = refl; val[worst] = fr;
        } else {
            std::vector<double> con = combine(centroid, simplex[worst], -0.5);
            const double fc = f(con); ++evals;
            if (fc > val[worst]) { simplex[worst] = con; val[worst] = fc; }
            else {
                for (size_t i = 1; i <= n; ++i) {
                    for (size_t d = 0; d < n; ++d) simplex[idx[i]][d] = simplex[best][d] + 0.5 * (simplex[idx[i]][d] - simplex[best][d]);
                    val[idx[i]] = f(simplex[idx[i]]); ++evals;
                }
            }
        }
    }
    size_t b = 0;
    for (size_t i = 1; i <= n; ++i) if (val[i] > val[b]) b = i;
    return simplex[b];
}

int main(int argc, char *argv[]) {
    if (argc < 5) {
        std::cout << "Not enough arguments" << std::endl;
        std::cout << "arguments: <GraphFile Path> <p value> <0 for getAngles> <file path to output angle file> [max objective evaluations]\n";
        return -1;
    }
    const int p = atoi(argv[2]), mode = atoi(argv[3]);
    mkdir("output", 0755);
    if (mode == 1) {
        // <graph> <p> 1 <input angle file> <output answer file>: final cut string for given angles (maxcut.cpp:258-287)
        if (argc < 6) { std::cout << "Not enough arguments" << std::endl; return -1; }
        try {
            std::ifstream inAngles(argv[4]);
            std::vector<double> bg;
            for (int i = 0; i < 2 * p; ++i) { double z = 0.0; inAngles >> z; bg.push_back(z); }
            maxcutGetFinalString(argv[1], p, {}, bg, argv[5]);
        } catch (std::exception &e) { std::cout << e.what() << std::endl; return -1; } catch (const char *m) { std::cout << m << std::endl; return -1; }
        return 0;
    }
    if (mode != 0 && mode != 2) { std::cout << "mode must be 0, 1 or 2" << std::endl; return -1; }
    const std::string outputPath(mode == 2 ? "tempAngles.txt" : argv[4]);
#ifdef QTB_HAVE_NLOPT
    const int maxEvals = (mode == 0 && argc > 5) ? atoi(argv[5]) : 0;          // 0: no cap, like the reference
#else
    const int maxEvals = (mode == 0 && argc > 5) ? atoi(argv[5]) : 200;
#endif
    try {
        Timer clock;
        clock.start();
        ExtraData data(p, argv[1]);
        data.outputFile = outputPath;
        const device::Job job = device::Job::FromEnvironment();
        const bool lead = job.rank == 0;
        QaoaObjective objective(data, job.rank, job.world, job.allreduce);
        if (lead) {
            std::cout << "Planned " << objective.OwnedEdges().size() << " of " << data.pairs.size() << " edge terms on rank 0 of " << job.world << " in "
                      << clock.getElapsed() << " seconds; " << objective.UnitsPerEvaluation() << " units and " << objective.LaunchesPerEvaluation()
                      << " kernel launch(es) per objective evaluation on this rank" << std::endl;
        }
        std::vector<double> start(2 * p);
        for (int i = 0; i < p; ++i) { start[i] = 0.392699; start[i + p] = 0.785399; }
        double bestSeen = -1.0;
        auto F_p = [&](const std::vector<double> &bg) {
            const double v = objective(bg);
            if (lead) {
                std::ofstream angles(outputPath);
                for (double a : bg) angles << a << " ";
            }
            bestSeen = std::max(bestSeen, v);
            return v;
        };
        Timer opt;
        opt.start();
        int evals = 0;
#ifdef QTB_HAVE_NLOPT
        // the reference's optimiser call, maxcut.cpp:211-213 (every rank runs the same deterministic COBYLA on the same
        // allreduced values); nlopt reports "roundoff-limited" by exception once the trust region collapses, as in the reference
        struct Ctx { decltype(F_p) *f; int *evals; } cctx{&F_p, &evals};
        nlopt::opt optimization(nlopt::LN_COBYLA, 2 * p);
        optimization.set_max_objective([](const std::vector<double> &x, std::vector<double> &, void *d) -> double {
            Ctx *c = static_cast<Ctx *>(d);
            ++*c->evals;
            return (*c->f)(x);
        }, &cctx);
        // The reference sets no stopping criterion: its runs end when NLopt detects round-off ("nlopt roundoff-limited"), which the
        // run-to-run noise of its randomly ordered contractions triggers after a few hundred evaluations.  This objective is
        // deterministic, so that never happens: stop where the reference's 6-digit angle file can no longer change.
        optimization.set_xtol_abs(1e-9);
        optimization.set_maxeval(maxEvals > 0 ? maxEvals : 20000);
        std::vector<double> best = start;
        try {
            best = optimization.optimize(start);
        } catch (std::runtime_error &error) {
            if (lead) std::cout << error.what() << std::endl;              // e.g. "nlopt roundoff-limited" (maxcut.cpp:217-221)
        }
        const double seconds = opt.getElapsed();
        if (lead) {
            std::cout << "Optimiser: NLopt LN_COBYLA" << std::endl;
#else
        const std::vector<double> best = nelderMeadMaximise(F_p, start, maxEvals, 0.1, evals);
        const double seconds = opt.getElapsed();
        if (lead) {
            std::cout << "Optimiser: Nelder-Mead (built without NLopt; the reference uses COBYLA)" << std::endl;
            std::ofstream angles(outputPath);
            for (double a : best) angles << a << " ";
            angles.close();
#endif
            std::cout.precision(12);
            std::cout << "F_p(start) evaluations: " << evals << ", best F_p = " << bestSeen << std::endl;
            std::cout.precision(6);
            std::cout << "Terms per second: " << evals * static_cast<double>(data.pairs.size()) / seconds << std::endl;
            std::cout << "Took " << clock.getElapsed() << " seconds" << std::endl;
        }
        if (mode == 2 && lead) {                      // angles, then the final string with them (maxcut.cpp:293-324)
            std::remove("tempAngles.txt");
            maxcutGetFinalString(argv[1], p, {}, best, argv[4]);
        }
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
        return -1;
    } catch (const char *msg) {
        std::cout << msg << std::endl;
        return -1;
    }
    return 0;
}
