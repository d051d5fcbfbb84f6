// qtorch_b200/host/Exceptions.h -- exception types of the qTorch API surface.
// Same class names and what() texts as /root/reference/src/Exceptions.h:24-58 (user code catches them by
// type and the CLI prints what()), plus DeviceUnavailable for the B200 engine, which has no CPU fallback.
#pragma once
#include <exception>
#include <string>

namespace qtorch {

#define QTORCH_DEFINE_EXCEPTION(NAME, TEXT)                                   \
    class NAME : public std::exception {                                      \
    public:                                                                   \
        const char *what() const noexcept override { return TEXT; }           \
    }

QTORCH_DEFINE_EXCEPTION(InvalidFile, "Invalid Input or Output File Path");
QTORCH_DEFINE_EXCEPTION(InvalidFileFormat, "Invalid File Format.");
QTORCH_DEFINE_EXCEPTION(InvalidTensorNetwork,
                        "QASM File Does Not Entangle All Qubits - Please Add 2 qubit gates or check the specified number of qubits");
QTORCH_DEFINE_EXCEPTION(ContractionFailure, "Contraction Failed.");
QTORCH_DEFINE_EXCEPTION(InvalidUserContractionSequence, "Invalid User Defined Contraction Sequence.");
QTORCH_DEFINE_EXCEPTION(InvalidFunctionInput, "Input to Function Invalid.");
QTORCH_DEFINE_EXCEPTION(NumWiresVsNodeRank, "Number of wires different than tensor rank.");
QTORCH_DEFINE_EXCEPTION(InvalidContractionMethod, "Invalid Contraction Method.");
QTORCH_DEFINE_EXCEPTION(QbbFailure,
                        "Quick BB Failure.\n Details:\n quickbb_64 Executable: ELF 64-bit LSB executable,\n x86-64, version 1 "
                        "(GNU/Linux), statically linked,\n for GNU/Linux 2.6.24, not stripped.\n Please check that your system is "
                        "Linux and meets the requirements to run this binary.\n Otherwise, use the simple stochastic contraction method.");

#undef QTORCH_DEFINE_EXCEPTION

// Raised when libqtorch_b200 cannot run (no B200 / CUDA error).  Carries the engine's own message.
class DeviceUnavailable : public std::exception {
public:
    explicit DeviceUnavailable(const std::string &why) : mWhy("qtorch_b200 device engine unavailable: " + why) {}
    const char *what() const noexcept override { return mWhy.c_str(); }
private:
    std::string mWhy;
};

}  // namespace qtorch
