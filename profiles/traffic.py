#!/usr/bin/env python
"""DRAM bytes per C-ABI call of one C2 step from `ncu -i rep --page raw --csv` (one `--set full` capture of
profiles/one_step.py c2, filter conv_shift|tc_kernel|obs_prep|clip_optim) -> profiles/r2_traffic.json.

  python profiles/traffic.py <raw.csv> <out.json>

Kernels are matched to calls by name and program order (profiles/README.md, reading guide): the
conv_shift_kernel launches are conv1..conv4 forward (conv4 in two passes), then conv4 (two passes),
conv3, conv2 input gradients; the conv_shift_wgrad_kernel launches conv4 (two passes), conv3, conv2,
conv1; the three long tc_kernel<128> launches the FC forward, input-gradient and weight-gradient GEMMs.
"""
import csv
import json
import sys


def main(raw, out):
    rows = list(csv.reader(open(raw, errors="replace")))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    units = rows[1]

    def val(r, key):
        v = float(r[col[key]].replace(",", ""))
        u = units[col[key]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
                    "s": 1e3}.get(u, 1)

    ks = []
    for r in rows[2:]:
        if len(r) <= col["dram__bytes_read.sum"]:
            continue
        ks.append((r[col["Kernel Name"]], val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                   val(r, "gpu__time_duration.sum")))
    shift = [k for k in ks if "conv_shift_kernel" in k[0]]
    wgrad = [k for k in ks if "conv_shift_wgrad_kernel" in k[0]]
    gemm = [k for k in ks if "tc_kernel<128" in k[0] and k[2] > 1.0]
    res = {"_source": "dram__bytes_read.sum + dram__bytes_write.sum of the main kernel(s) of each C-ABI call, one "
                      "`ncu --set full --clock-control none` capture of one C2 step (profiles/one_step.py, "
                      "profiles/r2_top_kernels_full.md, profiles/traffic.py); bytes per launch of the call."}
    for k in ks:
        if "obs_prep" in k[0]:
            res["a2c_obs_prep_stack"] = k[1]
        if "clip_optim" in k[0]:
            res["a2c_clip_optim_step|bytes|52256768"] = k[1]
    L = ["5120:2>16,80x80,k3s1", "5120:16>24,80x80,k3s1", "5120:24>32,80x80,k3s2", "5120:32>64,40x40,k3s2"]
    if len(shift) == 9 and len(wgrad) == 5 and len(gemm) == 3:
        res["a2c_conv2d_fwd|flops|" + L[0]] = shift[0][1]
        res["a2c_conv2d_fwd|flops|" + L[1]] = shift[1][1]
        res["a2c_conv2d_fwd|flops|" + L[2]] = shift[2][1]
        res["a2c_conv2d_fwd|flops|" + L[3]] = shift[3][1] + shift[4][1]
        res["a2c_conv2d_bwd_data|flops|" + L[3]] = shift[5][1] + shift[6][1]
        res["a2c_conv2d_bwd_data|flops|" + L[2]] = shift[7][1]
        res["a2c_conv2d_bwd_data|flops|" + L[1]] = shift[8][1]
        res["a2c_conv2d_bwd_weight|flops|" + L[3]] = wgrad[0][1] + wgrad[1][1]
        res["a2c_conv2d_bwd_weight|flops|" + L[2]] = wgrad[2][1]
        res["a2c_conv2d_bwd_weight|flops|" + L[1]] = wgrad[3][1]
        res["a2c_conv2d_bwd_weight|flops|" + L[0]] = wgrad[4][1]
        res["a2c_gemm|flops|5120x2000x25600"] = gemm[0][1]
        res["a2c_gemm|flops|5120x25600x2000"] = gemm[1][1]
        res["a2c_gemm|flops|2000x25600x5120"] = gemm[2][1]
    else:
        res["_error"] = "unexpected launch counts: %d shift, %d wgrad, %d gemm" % (len(shift), len(wgrad), len(gemm))
    res["_total_bytes_of_listed_kernels"] = sum(k[1] for k in ks)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
