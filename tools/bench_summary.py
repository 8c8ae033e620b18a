import json,sys
for l in open(sys.argv[1]):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l)
        r=d.get('roofline',{})
        print('imgs/s',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'frac',round(r.get('frac',0),3), 'fwdbwd',round(r.get('msda_fwd_bwd',{}).get('frac',0),3), {k:round(v*1e3,1) for k,v in r.get('kernel_ms',{}).items()})
        oc=d.get('other_configs_kernel_level')
        if oc: print({k:{kk:(round(vv["ms"]*1e3,1),round(vv.get("frac",0),3)) for kk,vv in v.items() if isinstance(vv,dict) and "ms" in vv} for k,v in oc.items()})
