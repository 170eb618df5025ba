import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:round(v,1) for k,v in list(d.get("kernel_ms",{}).items())[:5]})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
