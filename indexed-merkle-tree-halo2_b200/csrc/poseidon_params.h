// Host-side Poseidon parameter derivation (see poseidon_params.cpp).
#pragma once
#include "poseidon.cuh"
#include "poseidon_spec.cuh"

namespace imt {
// Fills *out with the T=3, R_F=8, R_P=57 BN254 parameter set in Montgomery form.
void poseidon_params_generate(PoseidonParams* out);
// Any instance Poseidon::<Fr, t, t-1>::new(r_f, r_p): fills SpecLayout{t, r_f, r_p}.total() elements. False when the
// Grain stream yields a singular matrix (never for the standard instances).
bool poseidon_spec_generate(unsigned t, unsigned r_f, unsigned r_p, Fr* out);
}  // namespace imt
