mkdir -p gpurun_out/r3s
for k in "24 10" "26 10" "22 10" "28 10" "24 8" "24 12" "26 12" "20 8"; do
  set -- $k
  VT_REFILL=$1 VT_TRI_ROUND=$2 timeout 200 python tools/frames_in_flight_probe.py --worlds 1,8 --steps 24 > gpurun_out/r3s/fif_$1_$2.jsonl 2> gpurun_out/r3s/fif_$1_$2.err
  echo "== refill $1 tri $2"; python -c "
import json
for l in open('gpurun_out/r3s/fif_$1_$2.jsonl'):
    d=json.loads(l)
    if d['frames_in_flight']==2: print(d['world'], d['ms_per_step'], d['Mrays_s'])"
done
