// qtorch_b200/apps/qtorch_main.cpp -- the `qtorch <script.inp>` front-end on the B200 engine.
// Same script keys, console/result-file lines and exit codes as /root/reference/src/main.cpp:42-309
// (keys: qasm measurement contractmethod quickbbseconds threads qbbonly readqbbresonly 64bit outputpath
// user-contract-seq).  New optional keys: ">int device N" selects the GPU, ">string qbbdir DIR/" redirects
// the lg.cnf / qbb.out / qbb-stats.out files (default output/), and for the line-graph methods
//   ">int slicewires N"  index-slice the network over N wires (4^N slices; host/Slicing.h), -1 = as few as give every
//                        rank a slice;   ">int lanes N"  plan replicas per GPU (default 2).
//   ">string measurementlist FILE"  every line of FILE names a measurement file; the circuit is reduced, ordered and
//                        compiled ONCE (host/PlanCache.h) and each measurement is evaluated by swapping the rank-1 caps
//                        and replaying the compiled plan (one result line per measurement).
// Multi-GPU: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK in the environment (e.g.
// `torchrun --no-python --nproc-per-node 8 qtorch script.inp`): the slices are dealt round-robin over the ranks, the
// partial sums meet in one NCCL allreduce, rank 0 alone prints and writes the result file.  With WORLD_SIZE > 1 and no
// slicewires key, -1 is assumed.
#include <sys/stat.h>
#include <fstream>
#include <iostream>

#include "../host/qtorch.hpp"

static std::string braced(double seconds) { return std::string(" { ") + std::to_string(seconds) + std::string(" } "); }

static void setDefaults(leviParser &p) {
    p.mapInt["quickbbseconds"] = 20;          // reference main.cpp:316-336
    p.mapInt["threads"] = 2;
    p.mapBool["qbbonly"] = false;
    p.mapBool["readqbbresonly"] = false;
    p.mapString["outputpath"] = "output/qtorch.out";
}

int main(int argc, char *argv[]) {
    if (argc < 2) {
        std::cout << "Usage:\nExecutable <input file>\n";
        return -1;
    }
    leviParser in;
    setDefaults(in);
    in.readInputFile(std::string(argv[1]));
    mkdir("output", 0755);
    if (in.mapInt.count("device")) setenv("QTORCH_DEVICE", std::to_string(in.mapInt["device"]).c_str(), 1);

    // one process per GPU: join the job (NCCL) before anything else; ranks other than 0 stay silent
    device::Job job;
    try {
        job = device::Job::FromEnvironment();
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
        return -1;
    }
    const bool lead = job.rank == 0;
    const bool sliced = in.mapInt.count("slicewires") > 0 || job.world > 1;
    const int sliceWires = in.mapInt.count("slicewires") ? in.mapInt["slicewires"] : -1;
    std::ofstream devNull;
    if (!lead) { devNull.open("/dev/null"); std::cout.rdbuf(devNull.rdbuf()); }
    auto barrier = [&job]() { if (job.world > 1) { double one = 1.0; job.allreduce(&one, 1); } };

    std::cout << "QASM file: " << in.mapString["qasm"] << "\n";
    std::cout << "Meas file: " << in.mapString["measurement"] << "\n";
    std::cout << "Output file: " << in.mapString["outputpath"] << "\n";
    std::ofstream result(lead ? in.mapString["outputpath"] : std::string("/dev/null"));
    if (!result) {
        std::cout << "Invalid Output File Path" << std::endl;
        return -1;
    }
    auto report = [&result](const std::exception &e) {
        std::cout << e.what() << std::endl;
        result << e.what() << std::endl;
    };

    std::shared_ptr<Network> net;
    try {
        net = std::make_shared<Network>(in.mapString["qasm"], in.mapString["measurement"]);
    } catch (std::exception &e) {
        report(e);
        return -1;
    }

    std::cout << "========Threading Info========" << std::endl;
    if (in.mapInt["threads"] <= 0 || in.mapInt["threads"] > std::pow(4, THRESH_RANK_THREAD)) {
        std::cout << "Invalid Number of Threads in Input File. If it is a large number, try reducing the number of threads. "
                     "Thread number set to 2."
                  << std::endl;
        in.mapInt["threads"] = 2;
    }
    std::cout << "Number of Threads set to: " << in.mapInt["threads"] << std::endl;
    net->SetNumThreads(in.mapInt["threads"]);      // accepted for script compatibility; the arithmetic runs on the GPU
    std::cout << "=====End of Threading Info====\n\n";

    Timer clock;
    clock.start();
    bool ok = false;
    // a sliced run walks the plan on the host only (no arithmetic); the device then runs the slices of that plan
    SlicedContraction::Tensors inputTensors;
    if (sliced) {
        for (int i = 0; i < net->GetNumOriginalNodes(); ++i) inputTensors.push_back(net->GetAllNodes()[i]->GetTensorVals());
        device::Engine::SetPlanOnly(true);
    }
    try {
        net->ReduceCircuit();
    } catch (std::exception &e) {
        report(e);
        return -1;
    }
    std::cout << "Throughout, time elapsed after reading in circuit is given in { curly brackets }. "
                 "Time starts after circuit has been read in.\n\n";
    std::cout << "Reduced circuit (removed 1- and 2-qub gates) " << braced(clock.getElapsed()) << "\n\n";

    const std::string method = in.mapString["contractmethod"];
    std::cout << "Contraction method: " << method << "\n";
    if (method == "linegraph-qbb" || method == "linegraph-minfill") {
        std::cout << "Contraction method: Linegraph / tree decomposition\n";
        LineGraph lg(net);
        const bool inProcessOrdering = method == "linegraph-minfill";      // addition: no external quickbb binary
        if (in.mapString.count("qbbdir")) lg.SetQBBOutDirectory(in.mapString["qbbdir"]);
        const bool sixtyFour = !(in.mapBool.count("64bit") && !in.mapBool["64bit"]);
        const bool onlyOrder = in.mapBool["qbbonly"], onlyContract = !onlyOrder && in.mapBool["readqbbresonly"];
        try {
            if (onlyOrder) {
                std::cout << "qbbonly=true. Only running qbb on linegraph, not doing contraction.\n";
                std::cout << "quickbbseconds set to: " << in.mapInt["quickbbseconds"] << std::endl;
                if (inProcessOrdering) std::cout << "In-process min-fill ordering, width " << lg.runMinFill() << std::endl;
                else if (lead) lg.runQuickBB(in.mapInt["quickbbseconds"], &clock, sixtyFour);
                std::cout << "QuickBB has been run. Set qbbonly=false and readqbbresonly=true to contract network. Exiting.\n";
                return 0;
            }
            if (onlyContract) {
                std::cout << "readqbbresonly=true. Attempting to read previous qbb result, and contracting network.\n";
            } else if (inProcessOrdering) {
                std::cout << "In-process min-fill ordering, width " << lg.runMinFill() << std::endl;
            } else {
                std::cout << "quickbbseconds set to: " << in.mapInt["quickbbseconds"] << std::endl;
                if (lead) lg.runQuickBB(in.mapInt["quickbbseconds"], &clock, sixtyFour);      // one writer; the others read its file
                barrier();
            }
            ok = lg.LGContract();
        } catch (std::exception &e) {
            report(e);
        }
        if (ok && !sliced && in.mapString.count("measurementlist")) {
            // the same circuit under many measurements: compile once, replay per measurement (the walk above fixed the
            // ordering file; its own result is reported first, as usual)
            try {
                std::shared_ptr<CompiledCircuit> cc = PlanCache::Get().Lookup(in.mapString["qasm"], lg.QBBOutFile(), true);
                std::ifstream list(in.mapString["measurementlist"]);
                std::string mfile;
                while (std::getline(list, mfile)) {
                    if (mfile.empty()) continue;
                    const std::complex<double> v = cc->EvaluateFile(mfile);
                    std::cout << "Result of Contraction [" << mfile << "]: " << v << std::endl;
                    result << "Result of Contraction [" << mfile << "]: " << v << std::endl;
                }
                std::cout << "Compiled plan: " << cc->Launches() << " kernel launch(es) per measurement, " << cc->Evaluations() << " measurements evaluated" << std::endl;
            } catch (std::exception &e) {
                report(e);
                ok = false;
            }
        }
        if (ok && sliced) {
            device::Engine::SetPlanOnly(false);
            try {
                SlicedContraction run(slicing::PlanOfNetwork(*net), sliceWires, job, in.mapInt.count("lanes") ? in.mapInt["lanes"] : 2);
                const slicing::SlicedPlan &sp = run.Sliced();
                std::cout << "Index slicing: " << sp.wires.size() << " wire(s) cut, " << sp.NumSlices() << " slices dealt over " << job.world
                          << " rank(s); " << sp.nInvariant << " of " << sp.plan.steps.size() << " steps are slice-invariant and run once; peak rank of a slice "
                          << sp.peakRank << "; units per slice " << static_cast<double>(sp.unitsPerSlice) << std::endl;
                const std::complex<double> value = run.Contract(inputTensors);
                std::cout << "Result of Contraction" << (onlyContract ? " (also printed to file)" : "") << ": " << value << std::endl;
                result << "Result of Contraction: " << value << std::endl;
                if (const char *full = std::getenv("QTORCH_PRINT_FULL")) {
                    if (std::atoi(full)) { std::cout.precision(17); std::cout << "@@value " << value.real() << " " << value.imag() << std::endl; std::cout.precision(6); }
                }
            } catch (std::exception &e) {
                report(e);
                ok = false;
            }
        } else if (ok) {
            std::cout << "Result of Contraction" << (onlyContract ? " (also printed to file)" : "") << ": " << net->GetFinalValue() << std::endl;
            result << "Result of Contraction: " << net->GetFinalValue() << std::endl;
        }
    } else if (sliced) {
        std::cout << "Index slicing / multi-GPU runs need a line-graph contraction method.\n";
        return -1;
    } else if (method == "simple-stoch" || method == "user-defined") {
        const bool haveSeq = method == "user-defined" && in.mapString.count("user-contract-seq");
        if (method == "user-defined" && !haveSeq)
            std::cout << "User contraction sequence file was not defined - contracting via simple stochastic" << std::endl;
        std::shared_ptr<Network> done;
        try {
            ContractionTools tools(net);
            done = haveSeq ? tools.ContractUserDefinedSequenceOfWires(in.mapString["user-contract-seq"]) : tools.Contract(Stochastic);
            std::cout << "Result of contraction:\n" << tools.GetFinalVal() << "\n";
            result << "Result of Contraction: " << tools.GetFinalVal() << std::endl;
        } catch (std::exception &e) {
            report(e);
            return -1;
        }
        ok = done != nullptr;
    } else {
        std::cout << "Error. 'contractmethod' bad option.\n";
        return -1;
    }

    if (!ok) {
        std::cout << "ERROR. ABORTING.\n" << braced(clock.getElapsed()) << "\n";
        result << "ERROR. ABORTING.\n" << braced(clock.getElapsed()) << "\n";
        return -1;
    }
    std::cout << "Number of floating point ops in full contraction: " << net->getNumFloatOps() << "\n";
    result << "Number of floating point ops in full contraction: " << net->getNumFloatOps() << "\n";
    std::cout << "Contraction complete. " << braced(clock.getElapsed()) << "\n";
    result << "Contraction complete. " << braced(clock.getElapsed()) << "\n";
    return 0;
}
