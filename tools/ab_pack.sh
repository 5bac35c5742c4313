cd "$(dirname "$0")/.."
for F in "" "--no-pack"; do
  echo "== $F"; python tools/maps_probe.py cfg4 --direct --h-rows 4 $F 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['maps_on'], d['max_abs_diff'], d['differing_px'])"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "whole_mosaic_fast_vs" 2>&1 | grep -E "cfg4 whole|passed|failed"
