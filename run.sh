#!/usr/bin/env bash
# The two GPU-bound stages of the reference pipeline with its own flags, paths and file formats:
#   clustering/code/run.sh:1-4         python cli.py cluster --feature_path=... --out_path=... --meta_path=...
#   subset_selection/code/run.sh:1-5   python cli.py run --shards_path=... --meta_path=... --out_path=...
# (the reference's top-level run.sh:1-5 calls the selection stage without ever running the clustering stage it reads
# from; both are here, in order).  Usage: bash run.sh [DATA_DIR] [extra --a.b.c=v flags for the selection stage]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
DATA="${1:-$HERE/data}"
shift || true
FEATURES="${FEATURES:-$DATA/features/shard-000000.pkl}"
CLUSTERS="${CLUSTERS:-$DATA/clusters/shard-000000.pkl}"
cd "$HERE"
[ -f acav100m_b200/libacav_b200.so ] || python -m acav100m_b200.build
python -m acav100m_b200.clustering.cli cluster --feature_path="$FEATURES" --out_path="$DATA/clusters" \
  --meta_path="$DATA/videos"
python -m acav100m_b200.subset_selection.cli run --shards_path="$CLUSTERS" \
  --meta_path="$DATA/videos" \
  --out_path="$DATA/output.csv" "$@"
