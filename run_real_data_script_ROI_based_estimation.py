#!/usr/bin/env python
"""CLI of the ROI-based MET2 estimation — same flags and output-folder naming as the reference's
run_real_data_script_ROI_based_estimation.py; FA search, ROI reductions and ROI fits run on the GPU."""
import argparse
import os
import time

import numpy as np
from tabulate import tabulate

from multicomponent_t2_toolbox_b200.motor.motor_recon_met2_real_data_ROI import motor_recon_met2_ROIs


# (flag, type, default, choices, required, help) — flags, types, defaults and choices of the reference's ROI script
_FLAGS = [
    ("--path_to_folder", str, None, None, True, "folder that holds the data, mask, ROIs and the output directory (trailing '/')"),
    ("--input", str, None, None, True, "4-D multi-echo NIfTI file inside the folder"),
    ("--mask", str, None, None, True, "3-D brain-mask NIfTI file inside the folder"),
    ("--ROIs", str, None, None, True, "3-D integer label NIfTI file inside the folder (0 = background)"),
    ("--minTE", float, None, None, True, "first echo time = echo spacing, ms"),
    ("--nTE", int, 32, None, True, "number of echoes"),
    ("--TR", float, None, None, True, "repetition time, ms"),
    ("--FA_method", str, "spline", ["spline", "brute-force"], True, "flip-angle search method"),
    ("--FA_smooth", str, "yes", ["yes", "no"], True, "Gaussian-smooth the data used for the flip-angle search"),
    ("--denoise", str, "None", ["TV", "NESMA", "None"], True, "pre-processing denoiser (NESMA runs on the GPU; TV is not provided)"),
    ("--reg_matrix", str, "I", ["I", "L1", "L2", "InvT2"], True, "Tikhonov matrix of the per-ROI X2 fit"),
    ("--myelin_T2_cutoff", float, 40, None, True, "upper T2 bound of the myelin-water compartment, ms"),
    ("--numcores", int, -1, None, False, "number of workers; here: number of GPUs, -1 = all visible GPUs"),
]


def build_parser():
    parser = argparse.ArgumentParser(description='Myelin Water Imaging')
    for flag, typ, default, choices, required, text in _FLAGS:
        kw = dict(type=typ, default=default, help=text)
        if required:
            kw["required"] = True
        if choices:
            kw["choices"] = choices
        parser.add_argument(flag, **kw)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    start_time = time.time()
    folder = args.path_to_folder
    path_to_save_data = folder + 'recon_all_X2' + '-' + args.reg_matrix + '_ROI-based/'
    table = [['1. Regularization matrix     ', args.reg_matrix], ['2. FA estimation method      ', args.FA_method],
             ['3. Smooth image for FA est.  ', args.FA_smooth], ['4. Denoising method          ', args.denoise],
             ['5. TR(ms)                    ', args.TR], ['6. Min. TE                   ', args.minTE],
             ['7. Number of TEs             ', args.nTE], ['8. Myelin T2-cutoff (ms)     ', args.myelin_T2_cutoff]]
    print('-------------------------------')
    print(tabulate(table, headers=['Selected options             ', '   ']))
    try:
        os.mkdir(path_to_save_data)
    except OSError:
        print('Warning: this folder already exists. Results will be overwritten')
    TE_array = np.array(args.minTE * np.arange(1, args.nTE + 1))
    motor_recon_met2_ROIs(TE_array, folder + args.input, folder + args.mask, folder + args.ROIs, path_to_save_data,
                          args.TR, args.reg_matrix, args.denoise, args.FA_method, args.FA_smooth, args.myelin_T2_cutoff,
                          args.numcores)
    print("--- %s seconds ---" % (time.time() - start_time))


if __name__ == "__main__":
    main()
