/*
 * dorktracer_debug.h -- entry points that exist ONLY in instrumented debug builds of libdorktracer.so
 * (make variant NAME=x DEFS=-DDT_TRAV_STATS / -DDT_TIMELINE).  The shipped library does not export them and nothing in the
 * product path calls them; tests/_trav_stats.py and tests/_timeline.py use them on the GPU box for the measurements logged
 * under profiles/.
 */
#ifndef DORKTRACER_DEBUG_H
#define DORKTRACER_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif
#ifdef DT_TRAV_STATS
/* traversal event counters: 0 nodes, 1 triangle tests, 2 shape visits, 3 BLAS entries, 4 leaf-box confirmations, 5 steps, 6 rays */
void dt_debug_stats(unsigned long long* out /* [8] */, int reset);
#endif
#ifdef DT_TIMELINE
/* per persistent warp: (kind << 32 | rays of the launch), start, queue drained, exit (ns of %globaltimer); returns the record count */
int dt_debug_timeline(unsigned long long* out /* 4 words per record */, int max_records);
/* per-ray state-machine steps, [closest | any-hit][min(63, steps / 8)] */
void dt_debug_steps_hist(unsigned int* out /* [128] */, int reset);
#endif
#ifdef __cplusplus
}
#endif
#endif /* DORKTRACER_DEBUG_H */
