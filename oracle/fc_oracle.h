/* oracle/fc_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Prototypes of the plain-C CPU restatement (fco_*) of the FieldCalculations hot path.
 * One prototype per entry of include/fcb200_api.inc; see that file for conventions and
 * for the reference file:line every entry follows. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load liboracle; the product never does.
 */
#ifndef FC_ORACLE_H
#define FC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define FC_FN(name, args) int fco_##name args;
#include "fcb200_api.inc"
#undef FC_FN

#ifdef __cplusplus
}
#endif

#endif /* FC_ORACLE_H */
