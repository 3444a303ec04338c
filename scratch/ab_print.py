import json, os, sys
d = json.loads(sys.stdin.read()); b = d["breakdown_last_step_rank0"]
print(os.environ.get("TAG", ""), "it/s %.1f" % d["value"], "busy %.2f" % b.get("search_sm_busy", 0),
      "search %.0f mstats %.0f estep %.0f mix %.0f of %.0f ms" % (b["ms_search"], b["ms_mstats"], b["ms_estep"], b["ms_mixture"], b["ms_total"]),
      "cand/s in search %.1fM" % (d["cand_scores_per_s_in_search"] / 1e6), "winner", d["winner_start"], "iters", d["em_iters_per_step"])
