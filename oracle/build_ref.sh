#!/bin/bash
# oracle/build_ref.sh -- build the UNMODIFIED reference hot path into oracle/_ref/ref_driver, so that the oracle's Dopri5
# controller (row a4, "parity unpinned") can be pinned against the reference's own Boost.Odeint build in one command:
#
#     BOOST_ROOT=/path/to/boost EIGEN_ROOT=/path/to/eigen3 [REFERENCE_ROOT=/root/reference] oracle/build_ref.sh
#     python -m pytest tests/test_reference_build.py -q          # diffs logL and (accepted, rejected) with the oracle
#
# BOOST_ROOT must contain boost/numeric/odeint.hpp, EIGEN_ROOT must contain Eigen/Dense (both header-only for this path).
# Sources are compiled where they lie under REFERENCE_ROOT with the reference's default release flags (CMakeLists.txt:25-29:
# -O3 -DNDEBUG, no -march, hence no FMA contraction beyond what -ffp-contract=off forbids); nothing is copied into the repo
# and every output goes to oracle/_ref/ (git-ignored).  Neither library exists in this repository's build image (probed:
# no odeint / Eigen directory anywhere on disk), so this recipe could not be run there: treat a compile error as a bug here.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
: "${BOOST_ROOT:?set BOOST_ROOT to a directory that contains boost/numeric/odeint.hpp}"
: "${EIGEN_ROOT:?set EIGEN_ROOT to a directory that contains Eigen/Dense}"
[ -f "$BOOST_ROOT/boost/numeric/odeint.hpp" ] || { echo "no boost/numeric/odeint.hpp under $BOOST_ROOT" >&2; exit 1; }
[ -f "$EIGEN_ROOT/Eigen/Dense" ] || { echo "no Eigen/Dense under $EIGEN_ROOT" >&2; exit 1; }
[ -d "$REF/src/model" ] || { echo "no reference tree at $REF" >&2; exit 1; }
OUT="$HERE/_ref"
mkdir -p "$OUT/obj"
# the library sources of the path (CMakeLists.txt:70-108 minus the older SIR model, the optimizers and the report writers)
SRCS=(
  src/sir_age_structured/Simulator.cpp
  src/sir_age_structured/SimulationResultProcessor.cpp
  src/sir_age_structured/solvers/Dopri5SolverStrategy.cpp
  src/sir_age_structured/caching/SimulationCache.cpp
  src/model/AgeSEPAIHRDModel.cpp
  src/model/AgeSEPAIHRDsimulator.cpp
  src/model/PieceWiseConstantNPIStrategy.cpp
  src/model/PiecewiseConstantParameterStrategy.cpp
  src/model/parameters/SEPAIHRDParameterManager.cpp
  src/model/objectives/SEPAIHRDObjectiveFunction.cpp
  src/utils/FileUtils.cpp
  src/utils/ReadContactMatrix.cpp
  src/utils/GetCalibrationData.cpp
  src/utils/ReadCalibrationConfiguration.cpp
  src/exceptions/CSVReadException.cpp
)
CXX="${CXX:-g++}"
FLAGS=(-std=c++17 -O3 -DNDEBUG -DEIGEN_NO_DEBUG -ffp-contract=off -fopenmp -I"$REF/include" -I"$REF/src" -isystem "$EIGEN_ROOT" -isystem "$BOOST_ROOT")
OBJS=()
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_' | sed 's/\.cpp$/.o/')"
  echo "  CXX $s"
  "$CXX" "${FLAGS[@]}" -c "$REF/$s" -o "$o"
  OBJS+=("$o")
done
echo "  CXX oracle/ref_driver.cpp"
"$CXX" "${FLAGS[@]}" "$HERE/ref_driver.cpp" "${OBJS[@]}" -o "$OUT/ref_driver"
echo "built $OUT/ref_driver"
