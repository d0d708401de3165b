/* Stub <numa.h> for building the UNMODIFIED reference offload engine
 * (/root/reference/kv_connectors/llmd_fs_backend/csrc/storage) in an image
 * without libnuma.  The reference uses exactly one libnuma symbol,
 * numa_set_preferred() (thread_pool.cpp:80), as a placement hint; a no-op keeps
 * behaviour identical apart from NUMA page placement.  Test infrastructure only. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
static inline void numa_set_preferred(int node) { (void)node; }
static inline int numa_available(void) { return -1; }
#ifdef __cplusplus
}
#endif
