"""Do queued host->device copies delay a device->host copy issued later on ANOTHER stream?  (scan_host_many: uploads of the
next items are enqueued ahead; the hit-list copy of the running item is enqueued when its counters have landed.)"""
import time, torch
dev = torch.device("cuda", 0)
h_in = [torch.empty(64 << 20, dtype=torch.uint8).pin_memory() for _ in range(6)]
d_in = [torch.empty(64 << 20, dtype=torch.uint8, device=dev) for _ in range(6)]
d_out = torch.empty(32 << 20, dtype=torch.uint8, device=dev)
h_out = torch.empty(32 << 20, dtype=torch.uint8).pin_memory()
def run(prio_a, prio_b, label):
    sa = torch.cuda.Stream(priority=prio_a); sb = torch.cuda.Stream(priority=prio_b)
    for rep in range(3):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); ea = torch.cuda.Event(enable_timing=True); eb = torch.cuda.Event(enable_timing=True)
        e0.record(sa)
        with torch.cuda.stream(sa):
            for h, d in zip(h_in, d_in): d.copy_(h, non_blocking=True)
            ea.record(sa)
        time.sleep(0.002)                       # the D2H is issued 2 ms later, while the second upload is in flight
        with torch.cuda.stream(sb):
            sb.wait_event(e0)
            h_out.copy_(d_out, non_blocking=True)
            eb.record(sb)
        torch.cuda.synchronize()
    print("%-40s 6 x 64 MiB H2D done at %.2f ms; 32 MiB D2H (issued ~2 ms in) done at %.2f ms" % (label, e0.elapsed_time(ea), e0.elapsed_time(eb)))
run(0, 0, "both default priority")
run(0, -1, "D2H stream high priority")
run(-1, -1, "both high priority")
print("asyncEngineCount", torch.cuda.get_device_properties(0).__getattribute__("multi_processor_count"))
