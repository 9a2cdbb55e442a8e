// Host-side Poseidon parameter derivation (see poseidon_params.cpp).
#pragma once
#include "poseidon.cuh"

namespace imt {
// Fills *out with the T=3, R_F=8, R_P=57 BN254 parameter set in Montgomery form.
void poseidon_params_generate(PoseidonParams* out);
}  // namespace imt
