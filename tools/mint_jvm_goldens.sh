#!/bin/bash
# Mints golden vectors from the REAL sdrtrunk Java classes (needs a JDK >= 17 and a built sdrtrunk: its classes and
# dependency jars).  Not runnable in this repository's build image (no JVM): this is the path by which oracle-vs-Java
# parity leaves "unpinned".
#
#   tools/mint_jvm_goldens.sh /path/to/sdrtrunk [class path of its dependency jars]
#
# 1. tools/jvm_goldens.py export   writes the inputs of tests/golden/*.npz as little-endian float32 files
# 2. javac + java OracleHarness    drives ComplexPolyphaseChannelizerM2, the output processors, ComplexFIRFilter2, the
#                                  decimators, the AGC, the FM demodulators, the DQPSK demodulators and FilterFactory.getTaps
# 3. tools/jvm_goldens.py import   packs what they produced into tests/golden_jvm/*.npz (same layout as tests/golden/)
# tests/test_jvm_goldens.py then compares the oracle and the CUDA path with them.
set -euo pipefail
SDRTRUNK=${1:?path to a built sdrtrunk checkout}
EXTRA_CP=${2:-}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
WORK=$ROOT/tests/golden_jvm
mkdir -p "$WORK/in" "$WORK/out" "$WORK/classes"
python "$ROOT/tools/jvm_goldens.py" export "$WORK/in"
CP="$SDRTRUNK/build/classes/java/main:$SDRTRUNK/build/resources/main:$(find "$SDRTRUNK/build" "$HOME/.gradle" -name '*.jar' 2>/dev/null | tr '\n' ':')$EXTRA_CP"
javac -d "$WORK/classes" -cp "$CP" "$ROOT/java/src/io/github/dsheirer/gpu/OracleHarness.java"
java -cp "$WORK/classes:$CP" io.github.dsheirer.gpu.OracleHarness "$WORK"
python "$ROOT/tools/jvm_goldens.py" import "$WORK"
echo "minted: $(ls "$WORK"/*.npz)"
