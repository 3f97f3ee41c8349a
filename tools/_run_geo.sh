for lk in 0 5 7; do for span in 3 0; do
  echo "### VSR_LATENCY_K=$lk VSR_STEAL_SPAN=$span"
  VSR_LATENCY_K=$lk VSR_STEAL_SPAN=$span timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-api 2>gpurun_out/geo.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))"
  grep "^rank" gpurun_out/geo.err | cut -c1-420
done; done
