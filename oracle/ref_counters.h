/* TEST INFRASTRUCTURE ONLY.  Event counters that oracle/Makefile splices into a TEMPORARY copy of the
 * reference's jly_goicp.cpp at build time (the copy is deleted after compiling; nothing of the reference is
 * stored in this repo).  Index: 0 InnerBnB calls (jly_goicp.cpp:297), 1 translation pops (:314),
 * 2 translation sub-cubes (:331; cube.point evals += Nd each), 3 rotation pops (:680),
 * 4 rotation cubes (:768), 5 ICP calls (:107). */
#ifndef GOICP_REF_COUNTERS_H
#define GOICP_REF_COUNTERS_H
extern long long goicp_ref_cnt[8];
#endif
