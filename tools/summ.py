import json,sys
for w in sys.argv[1:]:
    try: d=json.loads([l for l in open("gpurun_out/b_%s.json"%w) if l.startswith("{")][-1])
    except Exception as e: print(w,"ERR",e); continue
    print(w,"fps %.0f ms/step %.3f e2e %.0f seq %s frame_frac %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"].get("single_slot_sequential_fps"),d["roofline"]["frame"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    for k in d["roofline"]["kernels"]: print("   %s %.1f us/launch share %.2f %.0f GB/s frac %.2f"%(k["kernel"],k["ms_per_launch"]*1e3,k["share"],k["achieved_gbs"],k["frac"]))
