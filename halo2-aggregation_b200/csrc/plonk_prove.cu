// plonk_prove.cu — prover pipeline (placeholder until the pipeline lands)
#include "plonk.hpp"
struct ProverState {};
void h2a_prover_state_free(h2a_ctx*, ProverState* p) { delete p; }
