for geo in "" "4:640:4:0" "2:640:4:0" "16:320:4:0" "8:640:6:0"; do
  echo "### VSR_GEOMETRY=$geo"
  VSR_GEOMETRY=$geo timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/geo.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))"
  head -1 gpurun_out/geo.err | cut -c1-420
done
