// MetConstants.h -- the part of the reference header of the same name (src/mi_fieldcalc/MetConstants.h) that has
// exported symbols or that callers of the field functions need: the standard-level tables used by
// pressure2FlightLevel (MC.h:86-89) and the ICAO standard-atmosphere scalar conversions (MC.h:93-125).  Host code
// only (plain scalars: nothing here is worth a GPU).  The saturation-table helper class `ewt_calculator` is an
// implementation detail of the field functions and is not part of this header.
#ifndef MI_FIELDCALC_METCONSTANTS_H
#define MI_FIELDCALC_METCONSTANTS_H

namespace miutil {
namespace constants {

const float t0 = 273.15, r = 287., cp = 1004., p0 = 1000., g = 9.8; /* ref:39-46 */
const double ft_per_m = 3.2808399;                                   /* ref:51 */

/* standard pressure levels (hPa) and their flight levels (100 feet), ref:86-89 */
const int nLevelTable = 16;
const float pLevelTable[nLevelTable] = {1000, 925, 850, 800, 700, 500, 400, 300, 250, 200, 150, 100, 70, 50, 30, 10};
const float fLevelTable[nLevelTable] = {5, 25, 50, 65, 100, 185, 235, 300, 340, 385, 445, 530, 605, 675, 780, 1020};

/* ICAO standard atmosphere: pressure (hPa) <-> geopotential altitude (m), ref:93-107 */
double ICAO_geo_altitude_from_pressure(double pressure);
double ICAO_pressure_from_geo_altitude(double altitude);
/* geopotential altitude (m) <-> flight level (100 feet; rounded to 500 feet one way), ref:109-125 */
int FL_from_geo_altitude(double a);
double geo_altitude_from_FL(double fl);

} // namespace constants
} // namespace miutil

#endif // MI_FIELDCALC_METCONSTANTS_H
