for pool in 8388608 12582912 16777216 20000000; do
  echo "== default lib POOL=$pool"
  HYPERION_B200_POOL=$pool timeout 120 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 3 2>&1 | tail -1
done
echo "== POOL=8388608 tau=0.01"
HYPERION_B200_POOL=8388608 timeout 120 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 0.01 --iters 3 2>&1 | tail -1
echo "== POOL=16777216 tau=0.01"
HYPERION_B200_POOL=16777216 timeout 120 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 0.01 --iters 3 2>&1 | tail -1
