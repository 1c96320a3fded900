import json,sys
for w in sys.argv[1:]:
    try: d=json.loads([l for l in open("gpurun_out/b_%s.json"%w) if l.startswith("{")][-1])
    except Exception as e: print(w,"ERR",e); continue
    def show(tag, d):
        e2e = d.get("e2e", {}).get("value", float("nan"))
        sf = d["config"].get("single_frame")
        print(tag,"fps %.0f ms/step %.3f e2e %.0f frame_frac %.3f own %.3f"%(d["value"],d["ms_per_step"],e2e,d["roofline"]["frame"]["frac"],d["roofline"]["frame"].get("own_frac",float("nan"))), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
        if sf: print("   single frame: graph %.0f fps / %.1f us sync; launches %.0f fps / %.1f us"%(sf["graph"]["back_to_back_fps"],sf["graph"]["sync_latency_us_median"],sf["launches"]["back_to_back_fps"],sf["launches"]["sync_latency_us_median"]))
        for k in d["roofline"]["kernels"]: print("   %s %.1f us/launch share %.2f %.0f GB/s frac %.2f"%(k["kernel"],k["ms_per_launch"]*1e3,k["share"],k["achieved_gbs"],k["frac"]))
        if "nvlink" in d["roofline"]: print("   nvlink", d["roofline"]["nvlink"])
    show(w, d)
    if "e2e_packed_f16" in d: print("   e2e packed f16 %.0f"%d["e2e_packed_f16"]["value"])
    if "comparison" in d: print("   cufft", {k:v for k,v in d["comparison"]["cufft"].items() if k!="what"})
    if "cpu_baseline" in d: print("   cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    for c,v in (d.get("configs") or {}).items():
        if v is None or "error" in v: print("  ",c,v); continue
        show("  configs."+c, v)
