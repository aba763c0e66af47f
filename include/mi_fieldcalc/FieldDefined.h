/* mi_fieldcalc/FieldDefined.h -- drop-in replacement header (B200 build).
 *
 * Same names, values and signatures as the reference's src/mi_fieldcalc/FieldDefined.h:35-47, so that
 * code written against mi-fieldcalc compiles and links unchanged against libmi-fieldcalc.so.0 from
 * this repository (mi-fieldcalc_b200/csrc/shim.cc).
 */
#ifndef MI_FIELDCALC_FIELDDEFINED_H
#define MI_FIELDCALC_FIELDDEFINED_H

#include <cstdlib>

/* the undefined value callers conventionally use: 1.0e35f (reference FieldDefined.cc:34, 87) */
extern const float fieldUndef;

namespace miutil {

extern const float UNDEF;

/* tri-state "which values of this field are defined" flag; the numeric values are part of the ABI */
enum ValuesDefined { ALL_DEFINED = 0, NONE_DEFINED, SOME_DEFINED };

/* scan a HOST array: a value counts as defined when it is < UNDEF (reference FieldDefined.cc:36-60) */
ValuesDefined checkDefined(const float* data, size_t n);
/* flag from a count: 0 -> ALL, n -> NONE, else SOME (reference FieldDefined.cc:62-70) */
ValuesDefined checkDefined(size_t n_undefined, size_t n);
/* flag of a result that needs both a and b (reference FieldDefined.cc:72-83) */
ValuesDefined combineDefined(ValuesDefined a, ValuesDefined b);

} // namespace miutil

#endif // MI_FIELDCALC_FIELDDEFINED_H
