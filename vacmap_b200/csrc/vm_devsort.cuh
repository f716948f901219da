// Counting sort of small integer keys on the device (the fill planners' bucket sorts: jobs by slot class, band
// offset and length).  Order inside a bucket is unspecified -- members of a bucket are interchangeable for the
// planners (same class, nearly the same geometry), and what they plan is certified per job afterwards.
#pragma once
#include "vm_common.cuh"

// keys[n]: bucket in [0, n_key) or < 0 (left out).  After the call: start[n_key + 1] exclusive prefix sums (start[n_key] =
// number of live items), order[start[k] .. start[k + 1]) = the items of bucket k.  cursor[n_key] is scratch.
// Returns the number of kernel launches.
int vm_bucket_sort(const int32_t *keys, int n, int n_key, int32_t *start, int32_t *cursor, int32_t *order, cudaStream_t stream);
