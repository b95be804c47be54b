timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tile_staged or split or path_length" 2>&1 | tail -5
for cfg in "1 16777216 1000000" "1 16777216 300000" "1 8388608 1000000" "1 4194304 1000000"; do
  set -- $cfg
  echo "== TILES=$1 POOL=$2 TILE_MIN=$3"
  HYPERION_B200_TILES=$1 HYPERION_B200_POOL=$2 HYPERION_B200_TILE_MIN=$3 timeout 120 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 3 2>&1 | tail -1
done
HYPERION_B200_TIMING=1 HYPERION_B200_TILES=1 HYPERION_B200_POOL=16777216 timeout 120 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 2>&1 | grep "round [1-9]\]\|round 1[0-9]\]\|timing"
