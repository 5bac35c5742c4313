cd /root/repo
for B in 3 4; do
  P360_NVCC_DEFS="-DP360_TILE_BLOCKS=$B" python -m pano360_b200.build --force > /dev/null
  echo "== blocks=$B"; python tools/maps_probe.py cfg4 --direct --h-rows 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['maps_on'], d['max_abs_diff'], d['differing_px'])"
done
