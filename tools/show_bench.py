import json, sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(d['config']['workload'][:5], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'],2), 'Mpix/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], 'host_ms', round(d.get('host_ms_per_step',0),1), 'clk', d.get('clocks',{}).get('sm_mhz'), 'enqueue_ms', round(d.get('host_enqueue_ms_per_step',0),2))
        for k,v in d.get('kernels',{}).items(): print('   ', k, round(v['ms_per_step'],2), 'ms', round(v['GBps']), 'GB/s', 'n', v['launches_per_step'])
    elif any(w in line for w in ('passed','failed','Error','error','FAILED','assert')): print(line.strip()[:300])
