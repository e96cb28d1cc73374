#!/usr/bin/env python
"""CLI of the ROI-based MET2 estimation — same flags and output-folder naming as the reference's
run_real_data_script_ROI_based_estimation.py; FA search, ROI reductions and ROI fits run on the GPU."""
import argparse
import os
import time

import numpy as np
from tabulate import tabulate

from multicomponent_t2_toolbox_b200.motor.motor_recon_met2_real_data_ROI import motor_recon_met2_ROIs


def build_parser():
    parser = argparse.ArgumentParser(description='Myelin Water Imaging')
    parser.add_argument("--path_to_folder", default=None, type=str, help="Path to the folder where the data is located, e.g., /home/Datasets/MET2/", required=True)
    parser.add_argument("--input", default=None, type=str, help="Input data, e.g., Data.nii.gz", required=True)
    parser.add_argument("--mask", default=None, type=str, help="Brain mask, e.g., Mask.nii.gz", required=True)
    parser.add_argument("--ROIs", default=None, type=str, help="Brain ROIs, e.g., ROIs.nii.gz", required=True)
    parser.add_argument("--minTE", default=None, type=float, help="Minimum Echo Time (TE, units: ms)", required=True)
    parser.add_argument("--nTE", default=32, type=int, help="Number of TEs", required=True)
    parser.add_argument("--TR", default=None, type=float, help="Repetition Time (units: ms)", required=True)
    parser.add_argument("--FA_method", choices=["spline", "brute-force"], required=True, type=str, default="spline", help="Method to estimate the flip angle (FA)")
    parser.add_argument("--FA_smooth", choices=["yes", "no"], required=True, type=str, default="yes", help="Smooth data for estimating the FA")
    parser.add_argument("--denoise", choices=["TV", "NESMA", "None"], required=True, type=str, default="None", help="Denoising method")
    parser.add_argument("--reg_matrix", choices=["I", "L1", "L2", "InvT2"], required=True, type=str, default="I", help="Regularization matrix")
    parser.add_argument("--myelin_T2_cutoff", default=40, type=float, help="Maximum T2 for the myelin compartment: T2 threshold (units: ms)", required=True)
    parser.add_argument("--numcores", default=-1, type=int, help="Number of cores (ignored: the fit runs on the GPU)")
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
