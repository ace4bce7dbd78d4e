/* Test-infrastructure shim (NOT product code): the minimum of the Arduino core
 * that /root/reference/SRC/AudioSDRlib/AudioSDR.{h,cpp} touches, so that the
 * unmodified reference compiles on an x86-64 host.  See oracle/README.md.
 *   PI        - Arduino defines it as this double literal (reference H:208,249,359).
 *   boolean   - Arduino typedef.
 *   abs       - Arduino macro form; the reference needs abs(float) at C:141.
 *   Serial    - fast_sqrt_f32 prints from inside the hot path (H:444); swallowed. */
#ifndef ORACLE_SHIM_ARDUINO_H
#define ORACLE_SHIM_ARDUINO_H
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifndef PI
#define PI 3.1415926535897932384626433832795
#endif
typedef bool boolean;
#ifdef abs
#undef abs
#endif
#define abs(x) ((x) > 0 ? (x) : -(x))
struct OracleNullSerial {
  template <class... A> void print(A...) {}
  template <class... A> void println(A...) {}
  template <class... A> void begin(A...) {}
};
static OracleNullSerial Serial;
#endif
