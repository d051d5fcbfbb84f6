// oracle/stubs/nlopt.hpp -- TEST INFRASTRUCTURE ONLY.  Empty stand-in for NLopt's C++ header so that the
// reference's src/maxcut.h (which includes <nlopt.hpp> but whose graph / circuit helpers -- ExtraData,
// outputInitialPlusStateToFile, applyU_CsThenU_Bs -- never touch NLopt) compiles inside oracle/ref_harness.cpp.
// The optimiser itself (third-party COBYLA, /root/reference/nlopt-2.4.2) is out of scope.
#pragma once
