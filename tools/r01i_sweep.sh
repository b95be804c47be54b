for v in cb5 cb4 cb3; do
  for tau in 1 0.01; do
  echo "== SORT=2 $v tau=$tau"
  HYPERION_B200_SORT=2 HYPERION_B200_LIB=build/variants/libhyp_$v.so timeout 120 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau $tau --iters 3 2>&1 | tail -1
  done
done
