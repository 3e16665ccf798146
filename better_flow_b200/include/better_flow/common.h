// better_flow/common.h -- host-side mirror of the reference's common definitions
// (reference: better_flow_core/include/better_flow/common.h).  Same names, but the sensor size and
// the slice buffer limits are run-time values instead of compile-time macros, and OpenCV is gone:
// on the hot path the reference used cv:: only as a float container (see image.h).
#ifndef BF_COMMON_H
#define BF_COMMON_H

#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

typedef long int lint;
typedef long long int sll;
typedef unsigned int uint;
typedef unsigned long int ulong;
typedef unsigned long long int ull;

#define BF_VERSION "1.0-b200"

// Time conversion (common.h:35-36)
#define FROM_SEC(in) ull(1000000000 * (in))
#define FROM_MS(in) ull(1000000 * (in))

// Camera resolution: RES_X = sensor rows, RES_Y = sensor columns (common.h:39-40 hard-codes 180 / 240).
namespace bf {
struct SensorConfig {
    int rows = 180;
    int cols = 240;
};
inline SensorConfig &sensor() {
    static SensorConfig s;
    return s;
}
inline void set_sensor(int rows, int cols) {
    sensor().rows = rows;
    sensor().cols = cols;
}
}  // namespace bf
#define RES_X (bf::sensor().rows)
#define RES_Y (bf::sensor().cols)

#ifndef VERBOSE
#define VERBOSE false
#endif

// Z (time) component of the direction vector (common.h:60) and the timestamp divider (common.h:64)
#define NZ 127
#define T_DIVIDER 1

#include <better_flow/datastructures.h>

class Event;
typedef LinearEventCloudTemplate<Event> LinearEventCloud;
typedef LinearEventPtrsTemplate<Event> LinearEventPtrs;

#endif  // BF_COMMON_H
