"""Compact per-launch table from `ncu -i X.ncu-rep --page raw --csv` (the .ncu-rep files themselves stay on the GPU
box: a --set full capture of one register() is ~100 MB).  Usage: python tools/ncu_summary.py raw.csv > profiles/....txt"""
import csv
import re
import sys

COLS = [('gpu__time_duration.sum', 'time_ms', 1.0), ('dram__bytes_read.sum', 'dram_rd_MB', 1.0),
        ('dram__bytes_write.sum', 'dram_wr_MB', 1.0),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor%', 1.0),
        ('lts__t_sector_hit_rate.pct', 'L2hit%', 1.0),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%', 1.0),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts%', 1.0),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex%', 1.0),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%', 1.0),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%', 1.0),
        ('launch__registers_per_thread', 'regs', 1.0), ('launch__grid_size', 'grid', 1.0)]


def main(path):
  rows = list(csv.reader(open(path)))
  hdr, units, data = rows[0], rows[1], rows[2:]
  idx = {h: i for i, h in enumerate(hdr)}
  print('# ' + ' '.join(sys.argv))
  print('# units: time ms, dram bytes ' + units[idx['dram__bytes_read.sum']] +
        '; cold-cache, serialised launches (compare shares, not absolutes)')
  print('kernel'.ljust(44) + ' '.join(n.rjust(10) for _, n, _ in COLS))
  for r in data:
    name = re.sub(r'\(.*', '', r[idx['Kernel Name']]).replace('void ', '').replace('<unnamed>::', '')[:43]
    vals = []
    for key, _, _ in COLS:
      try:
        v = float(r[idx[key]].replace(',', ''))
        if key == 'gpu__time_duration.sum':      # ncu picks a unit per file: normalise to milliseconds
          v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(units[idx[key]], 1.0)
        vals.append('%.4g' % v)
      except Exception:   # noqa: BLE001
        vals.append('-')
    print(name.ljust(44) + ' '.join(v.rjust(10) for v in vals))


if __name__ == '__main__':
  main(sys.argv[1])
