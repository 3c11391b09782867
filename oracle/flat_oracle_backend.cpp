// oracle/flat_oracle_backend.cpp — TEST INFRASTRUCTURE.
// Gives the host-side drop-in (ahsoka_b200/host/alignmentstoreadset_b200.cpp) an
// ahs_phase_batch() that runs the CPU oracle, so that flatten + oracle + emission can be
// compared with the reference-verbatim binary on a machine without a GPU.  Linked ONLY into
// oracle/_ref/Ahsoka_flat_oracle; the product library never sees this file.
#include "../include/ahsoka_b200.h"
extern "C" int ahs_oracle_phase_batch(const ahs_batch_in*, ahs_batch_out*, int);
extern "C" void ahs_oracle_free_out(ahs_batch_out*);
extern "C" int ahs_phase_batch(const ahs_batch_in* in, ahs_batch_out* out, int) { return ahs_oracle_phase_batch(in, out, 1); }
extern "C" void ahs_free_out(ahs_batch_out* o) { ahs_oracle_free_out(o); }
extern "C" int ahs_warmup(int, uint64_t, uint64_t) { return 0; }
extern "C" const char* ahs_last_error(void) { return "oracle backend"; }
extern "C" int ahs_phase_batch_multi(const ahs_batch_in* in, ahs_batch_out* out, const int*, int) { return ahs_oracle_phase_batch(in, out, 1); }
