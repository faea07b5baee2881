"""Host-side enqueue time of one bench step (how far ahead of the GPU the launching thread runs)."""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    sd = bench.clip_state_dict(23)
    host = bench.build_host(4096, dev, sd)
    host.B = 1
    ids, mask, labels, fh, fv = bench.make_prompt(1)
    px = torch.randn(1, 100, 3, 224, 224, device=dev, dtype=torch.bfloat16)
    ins = [t.to(dev) for t in (ids, mask, labels, fh, fv)]
    hidden = torch.randn(1, bench.T_PROMPT + 355, 4096, device=dev, dtype=torch.bfloat16)

    def step():
        i, m, l, f, v = ins
        r = host.prepare_inputs_labels_for_multimodal(i, m, None, l, px, future_hands=f, future_valid=v, is_evaluate=False)
        return host.gather_hand_traj_states(hidden, r[4], future_valid=v, strict=False)

    with torch.no_grad():
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        t_host = (time.perf_counter() - t0) / 10
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) / 10
    print(f"host enqueue {t_host * 1e3:.2f} ms/step, wall {t_all * 1e3:.2f} ms/step")

    # host time by section (no synchronisation inside the timed calls)
    import cProfile, pstats, io
    pr = cProfile.Profile()
    with torch.no_grad():
        pr.enable()
        for _ in range(10):
            step()
        pr.disable()
    torch.cuda.synchronize()
    st = io.StringIO()
    pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(28)
    print(st.getvalue()[:6000])
