// libngsfhmm_b200_emulated.cpp - TEST INFRASTRUCTURE ONLY: the WHOLE product library - kernels, launchers and the
// C ABI layer nfh_ctx.cu - as one translation unit for the SIMT emulator (tests/simt/simt.h + simt_runtime.h), from
// the mechanically rewritten sources of tests/_simt_build.py.  tests/simt/run_gpu_suite_emulated.py builds it into a
// scratch directory together with the product's host library and command-line binary and runs the `-m gpu` tests'
// small cases against it: a pre-flight for the round-end GPU run on a machine without a GPU.  Never shipped, never
// loaded by the package.
#include "simt.h"

#include "nfh_estep.cu"
#include "nfh_lkl.cu"
#include "nfh_viterbi.cu"
// nfh_freq.cu is the second translation unit (libngsfhmm_b200_emulated_freq.cpp): its hundreds of kernel
// instantiations are most of the compile time, so the two units are compiled side by side
#include "nfh_ctx.cu"
