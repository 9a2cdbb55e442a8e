// Indexed-leaf logic: low-leaf lookups, non-inclusion witnesses, batched inserts.   (textually included by imt_capi.cu)
extern "C" imt_status imt_low_leaf_lookup(imt_tree* t, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched) {
    if (!t) return IMT_ERR_INVALID_ARG;
    (void)values; (void)q; (void)low_idx; (void)matched;
    return fail(t->ctx, IMT_ERR_INVALID_ARG, "not implemented yet");
}
extern "C" imt_status imt_non_inclusion_paths(imt_tree* t, const void* values, size_t q, uint64_t* low_idx, uint8_t* matched,
                                              void* low_leaves, void* siblings, uint8_t* helpers, uint8_t* is_largest) {
    if (!t) return IMT_ERR_INVALID_ARG;
    (void)values; (void)q; (void)low_idx; (void)matched; (void)low_leaves; (void)siblings; (void)helpers; (void)is_largest;
    return fail(t->ctx, IMT_ERR_INVALID_ARG, "not implemented yet");
}
extern "C" imt_status imt_insert_batch(imt_tree* t, const void* new_vals, size_t b, uint64_t first_idx, imt_insert_witness* w) {
    if (!t) return IMT_ERR_INVALID_ARG;
    (void)new_vals; (void)b; (void)first_idx; (void)w;
    return fail(t->ctx, IMT_ERR_INVALID_ARG, "not implemented yet");
}
