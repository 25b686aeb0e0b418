// Formulation (b): tcgen05 int8 MMA scan -- placeholder until the kernel lands (next commit).
#include "common.cuh"
#include "internal.h"

bool mma_supported(const smafa_db *) { return false; }
int mma_db_reserve(smafa_ctx *, smafa_db *, uint64_t) { return SMAFA_OK; }
int mma_db_pack(smafa_ctx *, smafa_db *, uint64_t, uint64_t) { return SMAFA_OK; }
void mma_db_free(smafa_db *db) { cudaFree(db->onehot); db->onehot = nullptr; }
int mma_scan(smafa_ctx *, const smafa_db *, smafa::ScanParams &, cudaStream_t) { return SMAFA_E_UNSUPPORTED; }
