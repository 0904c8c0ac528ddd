#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle). Builds the UNMODIFIED reference sources, where they lie
# under /root/reference, into oracle/_ref/ (git-ignored). Nothing is copied into the repo
# history; staged pub_*.h headers and objects live only under oracle/_ref/.
#
# Recipe follows SURVEY.md Appendix A (what src/main/CMakeLists.txt:174-201 does by hand):
#   1. stage pub_*.h as sleqp/<relpath>/pub_*.h, 2. hand-written export.h/defs.h,
#   3. gcc -std=gnu11 on the file list, 4. LAPACK shim onto SciPy's bundled OpenBLAS.
# Produces:
#   oracle/_ref/libsleqp_ref_lapack.so   reference core + reference fact_lapack.c  (dense oracle)
#   oracle/_ref/libsleqp_ref_b200.so     reference core + sleqp_b200/host/{fact/fact_b200.c, tr/tr_b200.c, sparse/mat_b200.c} (drop-in test)
#   oracle/_ref/eqp_harness_{lapack,b200,b200tr}   the EQP harness over the reference LAPACK backend, over our factorization
#                                        with the reference's Steihaug solver, and over our factorization + our TR solver
#   oracle/_ref/eqp_harness_b200aj       ... and with our augmented Jacobian (device KKT assembly) instead of standard_aug_jac.c
#   oracle/_ref/eqp_step_b200            the timing mode of the same harness (bench.py's reference-driven e2e leg)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(dirname "$HERE")"
REF="${SLEQP_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
SRC="$REF/src/main"
if [ ! -d "$SRC" ]; then echo "reference not present at $REF; keeping prebuilt $OUT" ; exit 0; fi
mkdir -p "$OUT/gen/sleqp" "$OUT/obj"

# 1. stage public headers
(cd "$SRC" && find . -name 'pub_*.h' | while read -r f; do
   mkdir -p "$OUT/gen/sleqp/$(dirname "$f")"; cp "$f" "$OUT/gen/sleqp/$f"; done)

# 2. generated headers
cat > "$OUT/gen/sleqp/export.h" <<'EOH'
#ifndef SLEQP_EXPORT_H
#define SLEQP_EXPORT_H
#define SLEQP_EXPORT
#define SLEQP_NO_EXPORT
#endif
EOH
cat > "$OUT/gen/sleqp/defs.h" <<'EOH'
#ifndef SLEQP_DEFS_H
#define SLEQP_DEFS_H
typedef enum { SLEQP_LP_SOLVER_HIGHS } SLEQP_LP_SOLVERS;
#define SLEQP_VERSION "1.0.2"
#define SLEQP_HAVE_ATTRIBUTE_WARN_UNUSED_RESULT
#define SLEQP_FORMAT_PRINTF(index, first)
#define SLEQP_VERSION_MAJOR 1
#define SLEQP_VERSION_MINOR 0
#define SLEQP_VERSION_PATCH 2
#define SLEQP_TRLIB_VERSION "absent"
#define SLEQP_GIT_BRANCH "oracle"
#define SLEQP_GIT_COMMIT_HASH "none"
#define SLEQP_LONG_VERSION "1.0.2 [oracle]"
#define SLEQP_LP_SOLVER SLEQP_LP_SOLVER_HIGHS
#define SLEQP_LP_SOLVER_NAME "none"
#define SLEQP_LP_SOLVER_VERSION "0"
#define SLEQP_LP_SOLVER_HIGHS_NAME "none"
#define SLEQP_LP_SOLVER_HIGHS_VERSION "0"
#define SLEQP_FACT_NAME "oracle"
#define SLEQP_FACT_VERSION "0"
#define SLEQP_FACT_LAPACK_NAME "LAPACK"
#define SLEQP_FACT_LAPACK_VERSION "scipy-openblas"
#define SLEQP_FACT_B200_NAME "B200"
#define SLEQP_FACT_B200_VERSION "0.1"
#endif
EOH
# newton.c includes <trlib.h> but uses no trlib symbol itself (newton.c:4,122)
echo "/* empty stub: trlib is absent in this environment */" > "$OUT/gen/trlib.h"

CORE="sparse/mat.c sparse/vec.c fact/fact.c aug_jac/aug_jac.c aug_jac/standard_aug_jac.c
 working_set.c iterate.c problem.c func.c settings.c timer.c log.c error.c cmp.c types.c enum.c
 util.c hess_struct.c dyn.c lsq.c tr/steihaug_solver.c tr/tr_solver.c tr/tr_util.c
 direction.c working_step.c scale.c problem_scaling.c feas.c"
CFLAGS="-std=gnu11 -O2 -fPIC -DNDEBUG -I$SRC -I$OUT/gen -I$OUT/gen/sleqp"
OBJS=""
for f in $CORE; do
  o="$OUT/obj/$(echo "$f" | tr '/' '_' | sed 's/\.c$/.o/')"
  gcc $CFLAGS -c "$SRC/$f" -o "$o"
  OBJS="$OBJS $o"
done

# 4. LAPACK shim onto SciPy's OpenBLAS (symbols are scipy_dgetrf_/scipy_dgetrs_)
OPENBLAS="$(python - <<'EOP'
import glob, os, scipy
d = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
print(sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so")))[0])
EOP
)"
cat > "$OUT/obj/lapack_shim.c" <<'EOC'
void scipy_dgetrf_(int*, int*, double*, int*, int*, int*);
void scipy_dgetrs_(char*, int*, int*, double*, int*, int*, double*, int*, int*);
void dgetrf_(int* M, int* N, double* A, int* LDA, int* IPIV, int* INFO)
{ scipy_dgetrf_(M, N, A, LDA, IPIV, INFO); }
void dgetrs_(char* T, int* N, int* NRHS, double* A, int* LDA, int* IPIV, double* B, int* LDB, int* INFO)
{ scipy_dgetrs_(T, N, NRHS, A, LDA, IPIV, B, LDB, INFO); }
EOC
gcc -O2 -fPIC -c "$OUT/obj/lapack_shim.c" -o "$OUT/obj/lapack_shim.o"
gcc $CFLAGS -c "$SRC/fact/fact_lapack.c" -o "$OUT/obj/fact_lapack.o"
gcc -shared -o "$OUT/libsleqp_ref_lapack.so" $OBJS "$OUT/obj/fact_lapack.o" "$OUT/obj/lapack_shim.o" \
    "$OPENBLAS" -Wl,-rpath,"$(dirname "$OPENBLAS")" -lm
echo "built $OUT/libsleqp_ref_lapack.so"
# EQP harness over the reference LAPACK backend (expected values of the SLEQP-iterate proxy)
gcc $CFLAGS -I"$SRC/fact" -I"$SRC/aug_jac" -I"$SRC/tr" "$HERE/eqp_harness.c" -o "$OUT/eqp_harness_lapack" \
    -L"$OUT" -lsleqp_ref_lapack -Wl,-rpath,'$ORIGIN' -lm

# drop-in: same reference core, our host glue instead of fact_lapack.c
HOST="$REPO/sleqp_b200/host"
if [ -f "$HOST/fact/fact_b200.c" ] && [ -f "$REPO/sleqp_b200/libsleqp_b200.so" ]; then
  GLUE=""
  for f in fact/fact_b200.c tr/tr_b200.c sparse/mat_b200.c aug_jac/b200_aug_jac.c; do
    o="$OUT/obj/$(basename "$f" .c).o"
    gcc $CFLAGS -I"$REPO/include" -I"$HOST" -I"$HOST/$(dirname "$f")" -I"$SRC/fact" -I"$SRC/tr" -I"$SRC/sparse" -I"$SRC/aug_jac" -c "$HOST/$f" -o "$o"
    GLUE="$GLUE $o"
  done
  gcc -shared -o "$OUT/libsleqp_ref_b200.so" $OBJS $GLUE \
      -L"$REPO/sleqp_b200" -lsleqp_b200 -Wl,-rpath,'$ORIGIN/../../sleqp_b200' -lm
  echo "built $OUT/libsleqp_ref_b200.so"
  HFLAGS="$CFLAGS -I$SRC/fact -I$SRC/aug_jac -I$SRC/tr -I$SRC/sparse -I$HOST -I$REPO/include"
  LINK="-L$OUT -lsleqp_ref_b200 -Wl,-rpath,\$ORIGIN -Wl,-rpath,\$ORIGIN/../../sleqp_b200 -L$REPO/sleqp_b200 -lsleqp_b200 -lm"
  gcc $HFLAGS "$HERE/eqp_harness.c" -o "$OUT/eqp_harness_b200" $LINK
  gcc $HFLAGS -DHARNESS_B200_TR "$HERE/eqp_harness.c" -o "$OUT/eqp_harness_b200tr" $LINK
  gcc $HFLAGS -DHARNESS_B200_TR -DHARNESS_B200_AUG_JAC "$HERE/eqp_harness.c" -o "$OUT/eqp_harness_b200aj" $LINK
  gcc $HFLAGS -DHARNESS_B200_TR -DHARNESS_B200_AUG_JAC -DHARNESS_TIMING "$HERE/eqp_harness.c" -o "$OUT/eqp_step_b200" $LINK
fi

# ---- the whole reference solver (sleqp_solver_solve) over an LP backend that exists here ---------------------------
# Every source of src/main except the backends whose libraries are absent (SuiteSparse, MUMPS, HSL, HiGHS, Gurobi,
# trlib); the LP backend is sleqp_b200/host/lp/lpi_simplex.c, the factorization either the
# reference's LAPACK backend or ours. sleqp_trlib_solver_create is a stub that raises (SLEQP_TR_SOLVER_CG is selected).
FULL=""
mkdir -p "$OUT/objfull"
for f in $(cd "$SRC" && find . -name '*.c' | grep -v -E 'fact_(cholmod|ma27|ma57|ma86|ma97|mumps|spqr|umfpack|lapack)|cholmod_helpers|hsl_matrix|lpi_gurobi|lpi_highs|trlib_solver|mpi_utils' | sort); do
  o="$OUT/objfull/$(echo "$f" | sed 's|^\./||' | tr '/' '_' | sed 's/\.c$/.o/')"
  if [ ! -f "$o" ] || [ "$SRC/$f" -nt "$o" ]; then gcc $CFLAGS -c "$SRC/$f" -o "$o"; fi
  FULL="$FULL $o"
done
cat > "$OUT/obj/trlib_stub.c" <<'EOC'
#include "tr/trlib_solver.h"
#include "error.h"
SLEQP_RETCODE
sleqp_trlib_solver_create(SleqpTRSolver** star, SleqpProblem* problem, SleqpSettings* settings)
{
  (void)star; (void)problem; (void)settings;
  sleqp_raise(SLEQP_INTERNAL_ERROR, "trlib is absent in this environment: select SLEQP_TR_SOLVER_CG");
}
EOC
gcc $CFLAGS -I"$SRC/tr" -c "$OUT/obj/trlib_stub.c" -o "$OUT/obj/trlib_stub.o"
gcc $CFLAGS -I"$HOST" -I"$SRC/lp" -c "$HOST/lp/lpi_simplex.c" -o "$OUT/obj/lpi_simplex.o"
gcc -shared -o "$OUT/libsleqp_full_lapack.so" $FULL "$OUT/obj/fact_lapack.o" "$OUT/obj/lapack_shim.o" "$OUT/obj/lpi_simplex.o" "$OUT/obj/trlib_stub.o" \
    "$OPENBLAS" -Wl,-rpath,"$(dirname "$OPENBLAS")" -lm
echo "built $OUT/libsleqp_full_lapack.so"
gcc $CFLAGS -I"$SRC/fact" -I"$SRC/aug_jac" -I"$SRC/tr" "$HERE/full_solve.c" -o "$OUT/full_solve_lapack" -L"$OUT" -lsleqp_full_lapack -Wl,-rpath,'$ORIGIN' -lm
if [ -f "$OUT/obj/fact_b200.o" ]; then
  # the reference tree as the integration patch leaves it: newton.c / trial_point.c select the B200 TR solver and augmented Jacobian
  mkdir -p "$OUT/patched/src/main" "$OUT/patched/cmake"
  cp "$SRC/newton.c" "$SRC/trial_point.c" "$OUT/patched/src/main/"
  cp "$REF/cmake/SearchFact.cmake" "$REF/cmake/SearchLPS.cmake" "$OUT/patched/cmake/"
  (cd "$OUT/patched" && patch -p1 -s < "$REPO/cmake/sleqp_b200_backend.patch")
  FULLB=""
  for o in $FULL; do case "$o" in *objfull/newton.o|*objfull/trial_point.o) ;; *) FULLB="$FULLB $o";; esac; done
  for f in newton trial_point; do
    gcc $CFLAGS -I"$HOST" -I"$REPO/include" -c "$OUT/patched/src/main/$f.c" -o "$OUT/obj/${f}_b200.o"
    FULLB="$FULLB $OUT/obj/${f}_b200.o"
  done
  gcc -shared -o "$OUT/libsleqp_full_b200.so" $FULLB $GLUE "$OUT/obj/lpi_simplex.o" "$OUT/obj/trlib_stub.o" \
      -L"$REPO/sleqp_b200" -lsleqp_b200 -Wl,-rpath,'$ORIGIN/../../sleqp_b200' -lm
  echo "built $OUT/libsleqp_full_b200.so"
  gcc $CFLAGS -I"$SRC/fact" -I"$SRC/aug_jac" -I"$SRC/tr" "$HERE/full_solve.c" -o "$OUT/full_solve_b200" -L"$OUT" -lsleqp_full_b200 -Wl,-rpath,'$ORIGIN' \
      -Wl,-rpath,'$ORIGIN/../../sleqp_b200' -L"$REPO/sleqp_b200" -lsleqp_b200 -lm
fi
