#!/bin/bash
# Side-by-side run against the REAL reference: iff a JDK exists, compile the reference's Java sources
# against its own jars (no Ant needed), index + lclassify a small synthetic corpus with it and with this library, and
# compare the .result files (tools/compare_results.py: scores within 1e-5 relative, labels / hit sets equal on
# >= 99.99 % of the reads, disagreements listed).  This is what turns "parity unpinned" into a pinned oracle; it
# cannot run in the build image (no java / javac) and never runs on the GPU test boxes.
#
#   tools/ref_lucene_run.sh <reference checkout> <workdir> [make_parity_corpus.py options]
set -euo pipefail
if ! command -v java >/dev/null || ! command -v javac >/dev/null; then
    echo "java/javac not found: the reference cannot run here (side-by-side parity stays unpinned; the oracle is pinned by tests/golden/lucene_*.json)" >&2
    exit 3
fi
REF=${1:?reference checkout}; WORK=${2:?work directory}; shift 2
HERE=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$WORK/classes"
# 1. the reference, compiled from its sources where they lie (TaxonDB only needs sqlite at run time)
find "$REF/src" -name '*.java' > "$WORK/sources.txt"
javac -nowarn -encoding UTF-8 -d "$WORK/classes" -cp "$REF/libs/*" @"$WORK/sources.txt"
# 2. corpus + the two configuration files (same keys as the shipped configuration.json)
python "$HERE/tools/make_parity_corpus.py" "$WORK" "$@"
# 3. the reference: BioSpectra i / lc (src/biospectra/BioSpectra.java:56-108)
CP="$WORK/classes:$REF/libs/*"
java -Xmx8g -cp "$CP" biospectra.BioSpectra i  -j "$WORK/conf_ref.json" -r "$WORK/ref"
java -Xmx8g -cp "$CP" biospectra.BioSpectra lc -j "$WORK/conf_ref.json" -in "$WORK/reads" -out "$WORK/out_ref"
# 4. this library through its CLI mirror of the same two modes
"$HERE/biospectra_b200/biospectra_gpu" index     -j "$WORK/conf_gpu.json" -r "$WORK/ref"
"$HERE/biospectra_b200/biospectra_gpu" lclassify -j "$WORK/conf_gpu.json" -in "$WORK/reads" -out "$WORK/out_gpu"
# 5. compare
python "$HERE/tools/compare_results.py" "$WORK/out_ref/sample.fa.result" "$WORK/out_gpu/sample.fa.result"
