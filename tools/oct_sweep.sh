for t in 512 1024; do echo "== threads $t"; ORBX_OCT_THREADS=$t python tools/bench_configs.py 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin)
for k,v in d.items():
    if isinstance(v,dict) and 'stage_ms_per_batch' in v: print(' ',k,round(v['frames_per_s_resident']),'single',round(v['single_frame_ms'],3),'quadtree',round(v['stage_ms_per_batch']['quadtree'],3))
"; done
