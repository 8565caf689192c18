/* TEST INFRASTRUCTURE — see hammlet_oracle_impl.h.  Instantiates the restatement for
 * real_t = float (suffix _f32) and real_t = double (suffix _f64). */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define HO_MAXK 64

#define REAL float
#define SFX f32
#define RLOG logf
#define REXP expf
#define RSQRT sqrtf
#define RFABS fabsf
#define RMAX FLT_MAX
#include "hammlet_oracle_impl.h"
#undef REAL
#undef SFX
#undef RLOG
#undef REXP
#undef RSQRT
#undef RFABS
#undef RMAX

#define REAL double
#define SFX f64
#define RLOG log
#define REXP exp
#define RSQRT sqrt
#define RFABS fabs
#define RMAX DBL_MAX
#include "hammlet_oracle_impl.h"

int ho_max_states(void) { return HO_MAXK; }
