/* TEST INFRASTRUCTURE — plain-C CPU oracle ("port") for the registration hot path.
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load the library built from this file.
 * See c_oracle_impl.inc for what is restated and the reference lines followed.
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -> oracle/_build/liboracle.so)
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#define REAL float
#define SFX(n) n##_f32
#define FLOOR floorf
#define COS cosf
#define SIN sinf
#define TANH tanhf
#include "c_oracle_impl.inc"
#undef REAL
#undef SFX
#undef FLOOR
#undef COS
#undef SIN
#undef TANH

#define REAL double
#define SFX(n) n##_f64
#define FLOOR floor
#define COS cos
#define SIN sin
#define TANH tanh
#include "c_oracle_impl.inc"

int orc_version(void) { return 1; }
