"""profiles/<tag>_kernel_tour_full.txt from gpurun_out/tour_full_raw.csv (ncu --page raw --csv)."""
import csv, sys
tag = sys.argv[1]
rows = list(csv.reader(open('gpurun_out/tour_full_raw.csv')))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_rd'), ('dram__bytes_write.sum', 'dram_wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64%'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('lts__t_sector_hit_rate.pct', 'l2hit%')]
names = dict(want)
def f(r, w):
    v, u = r[ix[w]], units[ix[w]]
    try: x = float(v.replace(',', ''))
    except ValueError: return v
    if 'time' in w: return f"{x * 1e3 if u == 'ms' else (x if u == 'us' else x / 1e3):9.1f}us"
    if 'bytes' in w: return f"{x * {'Gbyte': 1e3, 'Mbyte': 1, 'Kbyte': 1e-3, 'byte': 1e-6}[u]:9.1f}MB"
    return f"{x:7.1f}" if '%' in names[w] else f"{int(x):5d}"
with open(f'profiles/{tag}_kernel_tour_full.txt', 'w') as o:
    o.write("# ncu --set full --clock-control none, tools/kernel_tour.py (N=8192 d=10 ExpSquared for K1-K4 and the gradient; c2 GP N=1000 d=2 Matern32 for K5, 200 steps)\n")
    o.write("# per captured launch; cold-cache serialised replays (gpurun_out/tour_full_raw.csv)\n")
    o.write(f"{'kernel':44s} " + " ".join(f"{n:>11s}" for _, n in want) + "\n")
    for r in data:
        name = r[ix['Kernel Name']].replace('<unnamed>::', '').replace('void ', '')[:44]
        o.write(f"{name:44s} " + " ".join(f"{f(r, w):>11s}" for w, _ in want) + "\n")
print(open(f'profiles/{tag}_kernel_tour_full.txt').read())
