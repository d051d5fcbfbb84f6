// qtorch_b200/apps/qtorch_main.cpp -- the `qtorch <script.inp>` front-end on the B200 engine.
// Same script keys, console/result-file lines and exit codes as /root/reference/src/main.cpp:42-309
// (keys: qasm measurement contractmethod quickbbseconds threads qbbonly readqbbresonly 64bit outputpath
// user-contract-seq).  New optional keys: ">int device N" selects the GPU, ">string qbbdir DIR/" redirects
// the lg.cnf / qbb.out / qbb-stats.out files (default output/).
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

    std::cout << "QASM file: " << in.mapString["qasm"] << "\n";
    std::cout << "Meas file: " << in.mapString["measurement"] << "\n";
    std::cout << "Output file: " << in.mapString["outputpath"] << "\n";
    std::ofstream result(in.mapString["outputpath"]);
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
                else lg.runQuickBB(in.mapInt["quickbbseconds"], &clock, sixtyFour);
                std::cout << "QuickBB has been run. Set qbbonly=false and readqbbresonly=true to contract network. Exiting.\n";
                return 0;
            }
            if (onlyContract) {
                std::cout << "readqbbresonly=true. Attempting to read previous qbb result, and contracting network.\n";
            } else if (inProcessOrdering) {
                std::cout << "In-process min-fill ordering, width " << lg.runMinFill() << std::endl;
            } else {
                std::cout << "quickbbseconds set to: " << in.mapInt["quickbbseconds"] << std::endl;
                lg.runQuickBB(in.mapInt["quickbbseconds"], &clock, sixtyFour);
            }
            ok = lg.LGContract();
        } catch (std::exception &e) {
            report(e);
        }
        if (ok) {
            std::cout << "Result of Contraction" << (onlyContract ? " (also printed to file)" : "") << ": " << net->GetFinalValue() << std::endl;
            result << "Result of Contraction: " << net->GetFinalValue() << std::endl;
        }
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
