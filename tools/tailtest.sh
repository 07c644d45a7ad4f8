G=${1:-256}
for env in "AMGB_NO_TAIL=1" "AMGB_TAIL_NNZ=600000" "AMGB_TAIL_NNZ=600000 AMGB_TILE_MIN_NNZ=100000" "AMGB_TAIL_NNZ=600000 AMGB_TILE_MIN_NNZ=400000" "AMGB_TAIL_NNZ=600000 AMGB_TILE_MIN_NNZ=1500000" "AMGB_TAIL_NNZ=60000 AMGB_TILE_MIN_NNZ=400000"; do
  echo "== $env"; env $env AMGB_VERBOSE=1 python tools/tune_tiles.py --grid $G --settings "1,9,0" 2>&1 | grep -E "amgb\]|cycle_ms" | cut -c1-170
done
