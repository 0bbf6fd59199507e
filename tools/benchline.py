import json,sys
d=json.loads(open(sys.argv[1]).read())
print(sys.argv[2], "value %.1f" % d["value"], "passes", ["%.3f" % x for x in d["passes_ms_per_step"]], "flushed %.3f" % d["step_ms_flushed"]["median"], "e2e %.1f" % d["e2e"]["value"], "conv_ms %.3f" % d["roofline"]["kernel_ms"], "vote_ms %.3f" % d["roofline"]["vote"]["kernel_ms"])
