// placeholder: replaced below by the fused multi-GPU tridiagonalisation
#include "comm.h"
#include "internal.h"
namespace rs {
bool sytrd_mg_available(int64_t) { return false; }
void sytrd_mg_run(double*, int64_t, double*, double*, double*) { fail(1, __FILE__, __LINE__, "sytrd_mg: not built"); }
}  // namespace rs
