#!/usr/bin/env python
"""CLI of the voxel-wise MET2 reconstruction — same flags, defaults and output-folder naming as the reference's
run_real_data_script.py:18-62,82-93,119-122; the fit itself runs on the GPU (multicomponent_t2_toolbox_b200).
Plotting (--savefig) needs matplotlib + LaTeX, which this image lacks; the NIfTI outputs carry the same names the
reference's plot scripts read."""
from __future__ import division

import argparse
import os
import time

import numpy as np
from tabulate import tabulate

from multicomponent_t2_toolbox_b200.motor.motor_recon_met2_real_data import motor_recon_met2


def build_parser():
    parser = argparse.ArgumentParser(description='Myelin Water Imaging')
    parser.add_argument("--path_to_folder", default=None, type=str, help="Path to the folder where the data is located, e.g., /home/Datasets/MET2/", required=True)
    parser.add_argument("--input", default=None, type=str, help="Input data, e.g., Data.nii.gz", required=True)
    parser.add_argument("--mask", default=None, type=str, help="Brain mask, e.g., Mask.nii.gz", required=True)
    parser.add_argument("--minTE", default=None, type=float, help="Minimum Echo Time (TE, in ms)", required=True)
    parser.add_argument("--nTE", default=32, type=int, help="Number of TEs", required=True)
    parser.add_argument("--TR", default=None, type=float, help="Repetition Time (TR, in ms)", required=True)
    parser.add_argument("--FA_method", default='spline', type=str, help="Method to estimate the flip angle (FA)", choices=["spline", "brute-force"], required=True)
    parser.add_argument("--FA_smooth", default='yes', type=str, help="Smooth data for estimating the FA", choices=["yes", "no"], required=True)
    parser.add_argument("--denoise", default='TV', type=str, help="Denoising method", choices=["TV", "NESMA", "None"], required=True)
    parser.add_argument("--reg_method", default='X2', type=str, help="Regularization algorithm", choices=["NNLS", "T2SPARC", "X2", "L_curve", "GCV", "BayesReg"], required=True)
    parser.add_argument("--reg_matrix", default='I', type=str, help="Regularization matrix", choices=["I", "L1", "L2", "InvT2"], required=True)
    parser.add_argument("--numcores", default=-1, type=int, help="Number of cores used in the parallel processing (ignored: the fit runs on the GPU)", required=True)
    parser.add_argument("--myelin_T2_cutoff", default=40, type=float, help="Maximum T2 for the myelin compartment: T2 threshold (in ms)", required=True)
    parser.add_argument("--savefig", default='no', type=str, help="Save reconstructed maps in .png", choices=["yes", "no"], required=True)
    parser.add_argument("--savefig_slice", default=30, type=int, help="Axial slice to save reconstructed maps, e.g., --Slice=30", required=True)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    start_time = time.time()
    path_to_folder = args.path_to_folder
    path_to_data = path_to_folder + args.input
    path_to_mask = path_to_folder + args.mask
    reg_method, reg_matrix = args.reg_method, args.reg_matrix
    if reg_method in ('NNLS', 'T2SPARC'):
        path_to_save_data = path_to_folder + 'recon_all_' + reg_method + '/'
    else:
        path_to_save_data = path_to_folder + 'recon_all_' + reg_method + '-' + reg_matrix + '/'
    if reg_method == 'T2SPARC':
        reg_matrix = 'InvT2'
    headers = ['Selected options', '   ']
    table = [["Regularization method", reg_method], ["Regularization matrix", reg_matrix], ["Denoising method", args.denoise],
             ["TR (ms)", args.TR], ["Minimum TE (ms)", args.minTE], ["Number of TEs", args.nTE],
             ["Flip angle (FA) method", args.FA_method], ["Smooth image for FA estimation", args.FA_smooth],
             ["Myelin T2 cutoff (ms)", args.myelin_T2_cutoff], ["Number of cores (ignored, GPU)", args.numcores],
             ["Save figures", args.savefig]]
    print(tabulate(table, headers=headers, tablefmt="fancy_grid"))
    try:
        os.mkdir(path_to_save_data)
    except OSError:
        print("Creation of the directory %s failed" % path_to_save_data)
    else:
        print("Successfully created the directory %s " % path_to_save_data)
    TE_array = args.minTE * np.arange(1, args.nTE + 1)
    TE_array = np.array(TE_array)
    motor_recon_met2(TE_array, path_to_data, path_to_mask, path_to_save_data, args.TR, reg_method, reg_matrix,
                     args.denoise, args.FA_method, args.FA_smooth, args.myelin_T2_cutoff, args.numcores)
    if args.savefig == 'yes':
        print("--savefig: plotting needs matplotlib with LaTeX (not available here); the NIfTI maps were written to "
              + path_to_save_data)
    print("--- %s seconds ---" % (time.time() - start_time))


if __name__ == "__main__":
    main()
