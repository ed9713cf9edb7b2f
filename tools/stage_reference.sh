#!/bin/bash
# tools/stage_reference.sh -- stage the UNMODIFIED reference files the benchmark's reference legs and the
# script harness need into baseline/_ref/ (git-ignored, not gpurun-ignored: it travels to the GPU box,
# where /root/reference does not exist).  Nothing is edited; sources are copied byte for byte.
#   model.py, main_disentangled.py        the hot path and its only caller
#   data/cora/raw/ind.cora.*              Planetoid raw pickles            (BASELINE configs[0])
#   data_pre_false/chameleon/raw/*.npz    chameleon features + edges       (BASELINE configs[1])
#   data/squirrel/geom_gcn/raw/out1_graph_edges.txt   squirrel edge list (its features are not shipped)
#   load_data.py                          the reference's own twitch / fb100 parsers (reader fidelity tests)
#   data/citeseer/raw/ind.citeseer.*, data/{texas,wisconsin,cornell}/raw/out1_*.txt,
#   data/facebook100/{Amherst41,JohnsHopkins55,Reed98}.mat, data/twitch/{DE,PTBR}/musae_*, mini/year9.pt
#                                         the other datasets of hyperparameters_setting whose files are shipped
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
cp -f "$REF/load_data.py" $OUT/
mkdir -p $OUT/data/citeseer/raw $OUT/data/facebook100 $OUT/mini
cp -f "$REF"/data/citeseer/raw/ind.citeseer.* $OUT/data/citeseer/raw/
for n in texas wisconsin cornell; do
  mkdir -p $OUT/data/$n/raw
  cp -f "$REF"/data/$n/raw/out1_graph_edges.txt "$REF"/data/$n/raw/out1_node_feature_label.txt $OUT/data/$n/raw/
done
for n in Amherst41 JohnsHopkins55 Reed98; do cp -f "$REF/data/facebook100/$n.mat" $OUT/data/facebook100/; done
for l in DE PTBR; do mkdir -p $OUT/data/twitch/$l; cp -f "$REF"/data/twitch/$l/musae_${l}_* $OUT/data/twitch/$l/; done
cp -f "$REF/mini/year9.pt" $OUT/mini/
( cd $OUT && sha256sum model.py main_disentangled.py load_data.py > SHA256SUMS )
echo "staged $(find $OUT -type f | wc -l) files into $OUT ($(du -sh $OUT | cut -f1))"
