// ba_schur_tc.cu — per-track Schur complement on the 5th-generation tensor cores (tcgen05 / TMEM). Placeholder
// until the kernel lands: reports "does not apply" so that the SIMT kernel runs.
#include "ba_internal.h"

namespace ba {
int schur_tc_prepare_device() { return BA_OK; }
}  // namespace ba
