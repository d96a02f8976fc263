for lib in gsdf_slam_b200/libgsr.so variants/libgsr_adam128.so; do for c in 1 2; do
GSR_LIB_PATH=$PWD/$lib GSR_ADAM_SIDE_CTAS=$c python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib ctas',$c,'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d['clocks'], {k:round(v['ms']*1000,1) for k,v in d['stages'].items() if 'adam' in k or 'blend' in k})
"
GSR_LIB_PATH=$PWD/$lib GSR_ADAM_SIDE_CTAS=$c python tools/overlap_timeline.py | tail -1
done; done
