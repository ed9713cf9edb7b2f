#!/bin/bash
# tools/stage_reference.sh -- stage the UNMODIFIED reference files the benchmark's reference legs and the
# script harness need into baseline/_ref/ (git-ignored, not gpurun-ignored: it travels to the GPU box,
# where /root/reference does not exist).  Nothing is edited; sources are copied byte for byte.
#   model.py, main_disentangled.py        the hot path and its only caller
#   data/cora/raw/ind.cora.*              Planetoid raw pickles            (BASELINE configs[0])
#   data_pre_false/chameleon/raw/*.npz    chameleon features + edges       (BASELINE configs[1])
#   data/squirrel/geom_gcn/raw/out1_graph_edges.txt   squirrel edge list (its features are not shipped)
set -e
REF=${1:-/root/reference}
cd "$(dirname "$0")/.."
if [ ! -f "$REF/model.py" ]; then echo "stage_reference: $REF/model.py not found (nothing staged)"; exit 0; fi
OUT=baseline/_ref
mkdir -p $OUT/data/cora/raw $OUT/data_pre_false/chameleon/raw $OUT/data/squirrel/geom_gcn/raw
cp -f "$REF/model.py" "$REF/main_disentangled.py" $OUT/
cp -f "$REF"/data/cora/raw/ind.cora.* $OUT/data/cora/raw/
cp -f "$REF/data_pre_false/chameleon/raw/chameleon.npz" $OUT/data_pre_false/chameleon/raw/
cp -f "$REF/data/squirrel/geom_gcn/raw/out1_graph_edges.txt" $OUT/data/squirrel/geom_gcn/raw/
( cd $OUT && sha256sum model.py main_disentangled.py > SHA256SUMS )
echo "staged $(find $OUT -type f | wc -l) files into $OUT ($(du -sh $OUT | cut -f1))"
