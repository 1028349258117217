"""Print the interesting fields of a bench.py JSON line.  usage: show_bench.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.3e e2e %.3e ms/step %.3f e2e ms %.3f p50 %.3f p95 %.3f launches %d" % (d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['latency_ms_p50'], d['e2e']['latency_ms_p95'], d['gpu_launches']))
r = d['roofline']; print("roofline frac %.3f whole %.3f traffic %s stages %s" % (r['frac'], r['whole_step_frac'], r.get('traffic'), {k: round(v, 4) for k, v in r['stage_ms_per_step'].items()}))
print("tail", d.get('tail')); print("clocks", d['clocks'])
if 'cfg2_n1e8' in d:
    c = d['cfg2_n1e8']; print("cfg2: value %.3e ms %.3f e2e %.3e (%.3f ms) frac %.3f whole %.3f stream ms %.3f parity %s" % (c['value'], c['ms_per_step'], c['e2e_value'], c['e2e_ms_per_step'], c['roofline']['frac'], c['roofline']['whole_step_frac'], c['roofline']['kernel_ms_per_launch'], c['parity_vs_oracle_whole_series']))
    print("cpu", d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['parity_vs_oracle_last_query'])
    print("query_set", d['query_set'])
    for e, b in d['cnsm_dtw']['eps'].items():
        print("dtw eps", e, "queries", b['queries'], "value %.3e p50 %.2f dtws %d cells %.3e frac %s" % (b['value'], b['kernel_ms_p50'], b['dtws'], b['cells_executed'], b['roofline']['frac']))
        for r in b['rows']: print("   ", r['offset'], "%.2f" % r['kernel_ms'], [round(x, 2) for x in r['stage_ms']], r['gate_pass'], r['dtws'], r['answers'])
    print(d['window_mean'])
