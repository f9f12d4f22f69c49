#!/usr/bin/env python
"""NanoReviser_train on the B200 path: the command line of the reference's NanoReviser_train.py (same options, defaults, output
files and messages), with label generation in nanoreviser_b200/trainprep.py (re-segmentation on the GPU) and the two fits in
nanoreviser_b200/train.py (hand-written CUDA training operators, include/nrv_train.h).

    python NanoReviser_train.py -d <fast5 dir> -r <reference.fasta> -S <species> -M ./model/ -m <graphmap> [-e 50 -b 512 -w 13]

Flow (NanoReviser_train.py:124-221 of the reference): per read  get_read_data -> read fasta -> mapper -> SAM -> alignment columns
-> clipped / cleaned labels -> per-read .npz under <model_dir>/<species>/training_input/ ; then all .npz -> window tensors ->
model 1 fit -> weights + history csv + parameter json -> model 2 fit -> the same.

Extensions (both optional): --sam_dir <dir> takes <read name>.sam files that already exist instead of running the mapper
(the aligner is an external program; it is not part of this repository); --device picks the GPU.
"""
import csv
import json
import os
import shutil
import sys
import time
from optparse import OptionParser

import numpy as np


def get_args():
    p = OptionParser()
    p.add_option('-d', '--fast5_base_dir', action='store', type="string", default='./unitest/training_data/fast5/', dest='fast5_base_dir',
                 help='path to the fast5 files')
    p.add_option('-o', '--output_dir', action='store', type="string", dest='output_dir', default='./unitest/nanorev_training_result/',
                 help='path to store the output sumarry files')
    p.add_option('-r', '--reference', action='store', type="string", dest='genome_fn', default='./unitest/training_data/reference.fasta',
                 help='path to store the output files')
    p.add_option('--model_type', action='store', type="string", dest='model_type', default='both', help='both, model1 or model2, default is both')
    p.add_option('-S', '--species', action='store', type="string", dest='species', default='unitest',
                 help='species of training data, default is unitest for test Nanoreviser_train')
    p.add_option('-M', '--output_model', action='store', type="string", dest='model_dir', default='./model/', help='path to store the model files')
    p.add_option('-m', '--mapper_exe', action='store', type="string", dest='graphmap_exe', default='graphmap',
                 help='the align tool for generate the lable of training data, default is graphmap')
    p.add_option('-L', '--output_format', action='store', type="string", dest='output_format', default='sam')
    p.add_option("--thread", action="store", type="int", dest="thread", default=1, help='thread, default is 1')
    p.add_option('-t', '--tmp_dir', action='store', type="string", dest='temp_dir', default='./train_tmp/',
                 help='path to the tmp dir, which is used to store the preprocessing files')
    p.add_option('-f', '--failed_read', action='store', type="string", dest='failed_reads_filename', default='failed_reads.txt',
                 help='document to log the failed reads, default is failed_read.txt')
    p.add_option('-g', '--basecall_group', action='store', type="string", dest='basecall_group', default='Basecall_1D_000',
                 help='attrs for finding the events file in fast5 file, default is Basecall_1D_000')
    p.add_option('-s', '--basecall_subgroup', action='store', type="string", dest='basecall_subgroup', default='BaseCalled_template',
                 help='attrs for finding the events file in fast5 file, default is BaseCalled_template')
    p.add_option("-b", "--batch_size", action="store", type="int", dest="batch_size", default=512, help='batch size, default is 256')
    p.add_option("-e", "--epochs", action="store", type="int", dest="epochs", default=50, help='epochs, default is 50')
    p.add_option("-w", "--window_size", action="store", type="int", dest="window_size", default=13, help='window size, default is 13')
    p.add_option("-c", "--read_counts", action="store", type="int", dest="read_counts", default=0,
                 help='the number of read included in the training data, must smaller than the number of files stored in fast5 dir, '
                      'default is 0, which means use all the files')
    p.add_option("--validation_split", action="store", type="float", dest="validation_split", default=0.01,
                 help='validation data size, default is 0.01, which means 1% of the data are used as validation data')
    p.add_option('--model1_train_dir', action='store', type="string", dest='model1_train_dir', default='',
                 help='model dirs for trained model1, for transfer learning')
    p.add_option('--model2_train_dir', action='store', type="string", dest='model2_train_dir', default='',
                 help='model dirs for trained model2, for transfer learning')
    p.add_option('--test_mode', action='store_true', default=False, help='just for unitest')
    p.add_option("-v", "--virsion", action="store_true", dest="virsion", help="version of NanoReviser")
    # extensions
    p.add_option('--sam_dir', action='store', type="string", dest='sam_dir', default='',
                 help='directory of existing <read name>.sam files (skips running the mapper)')
    p.add_option('--device', action='store', type="int", dest='device', default=0, help='CUDA device')
    args, _ = p.parse_args()
    args.model_dir = str(args.model_dir) + '/' + str(args.species) + '/'
    args.train_input_dir = str(args.model_dir) + '/training_input/'
    args.train_model_dir = str(args.model_dir) + '/training_model/'
    if args.virsion:
        print("The virsion of NanoReviser : 1.0 ")
    return args


def check_path(path):
    if not os.path.exists(str(path)):
        try:
            os.makedirs(str(path))
        except Exception as e:
            raise FileNotFoundError('！！！[Error] make dir ' + str(path) + str(e))


def model_fn_generate(args, model_tag):
    """nanorevutils/fileoptions.py:57-75"""
    sg = str(args.species) + '_win' + str(args.window_size) + '_' + str(args.epochs) + 'ep_' + str(model_tag)
    return (args.model_dir + sg + '.h5', args.train_model_dir + 'train_' + sg + '.h5', args.output_dir + sg + '_hisroty.csv',
            args.output_dir + sg + '_parameters.json')


def summary_generate(args, start_t):
    return {'model_type': args.model_type, 'species': args.species, 'input_file': args.fast5_base_dir, 'read_counts': args.read_counts,
            'window_size': args.window_size, 'epochs': args.epochs, 'batch_size': args.batch_size, 'validation_split': args.validation_split,
            'training_time': str(int(time.time() - start_t)) + ' seconds'}


def write_sumery_file(history, summary, history_fn, summary_fn):
    try:
        with open(summary_fn, 'w') as f:
            json.dump(summary, f)
    except Exception as e:
        raise RuntimeError('！！！[Error] saveing summary hisroty result ', e)
    try:
        keys = list(history)
        with open(history_fn, 'w', newline='') as f:
            wr = csv.writer(f)
            wr.writerow(keys)
            for row in zip(*[history[k] for k in keys]):
                wr.writerow(row)
    except Exception as e:
        raise RuntimeError('！！！[Error] saveing training hisroty result ', e)


def handel_input_fast5(fast5_fn_sg, args, genome_index, reviser):
    """One read -> its training .npz (nanorevtrainutils.py:56-130)."""
    from nanoreviser_b200 import fast5, trainprep
    fast5_fn = os.path.join(args.fast5_base_dir, fast5_fn_sg)
    stem = fast5_fn_sg.split('.')[0]
    try:
        read = fast5.read_fast5_arrays(fast5_fn, args.basecall_group, args.basecall_subgroup)
    except Exception as e:
        raise EOFError('！！！[Error] ' + stem + str(e))
    try:
        if args.sam_dir:
            with open(os.path.join(args.sam_dir, stem + '.sam')) as fp:
                sam = trainprep.load_sam_record(fp.readlines())
        else:
            read_fasta_fn = args.temp_dir + stem + '.fasta'
            with open(read_fasta_fn, 'w') as fp:
                fp.write(">" + fast5_fn.replace(' ', '|||') + '\n' + read.bases.tobytes().decode('ascii') + '\n')
            if not args.test_mode:
                print('[p:::] ' + stem + '.fasta was saved for mapping......')
            out_fn = args.temp_dir + stem + '.sam'
            sam = trainprep.align_to_genome(out_fn, args.graphmap_exe,
                                            trainprep.prep_graphmap_options(args.genome_fn, read_fasta_fn, out_fn, args.output_format, 1))
            os.remove(out_fn)
            os.remove(read_fasta_fn)
        if not args.test_mode:
            print('[p:::] ' + stem + '.sam has been loaded......')
    except Exception as e:
        raise RuntimeError('！！！[Error] ' + stem + str(e))
    try:
        arrays = trainprep.read_training_arrays(read, sam, genome_index, reviser)
        np.savez(str(args.train_input_dir) + stem, **arrays)
        if not args.test_mode:
            print('[s:::] ' + stem + '.npz has been saved......')
    except Exception as e:
        raise RuntimeError('[！！！Error] ' + stem + str(e))


def train_preprocessing(args):
    from nanoreviser_b200 import engine, train, trainprep
    genome_index = trainprep.parse_fasta(args.genome_fn)
    if not args.test_mode:
        print(args.genome_fn, 'has been load......')
    fns = os.listdir(args.fast5_base_dir)
    if args.read_counts and args.read_counts < len(fns):
        fns = fns[:int(args.read_counts)]
    # the segmentation kernels belong to a handle; any weight set will do for it (only nrv_segment is used here)
    m1, m2 = train.init_weights(11, 6), train.init_weights(11, 5)
    with engine.Reviser(m1, m2, device=args.device) as rv:
        for fn in fns:
            handel_input_fast5(fn, args, genome_index, rv)


def fit_and_save(args, tag, n_class, signal_x, x, y, transfer_fn):
    from nanoreviser_b200 import train, weights
    start_t = time.time()
    w0 = weights.load_model_weights(transfer_fn) if transfer_fn else None
    model = train.TrainModel(window=args.window_size, n_class=n_class, weights=w0, device=args.device)
    history = model.fit([signal_x[:, :, :, np.newaxis], x, y], [y, np.zeros((len(y), 1))],
                        class_weight={0: 3, 1: 5, 2: 1, 3: 1, 4: 1, 5: 1}, validation_split=args.validation_split, shuffle=True,
                        epochs=args.epochs, batch_size=args.batch_size, verbose=0 if args.test_mode else 2)
    pre_fn, train_fn, history_fn, summary_fn = model_fn_generate(args, tag)
    model.save_weights(train_fn)            # (the train model's file additionally holds the centre embedding in Keras; here the predict layout)
    model.save_weights(pre_fn)
    write_sumery_file(history, summary_generate(args, start_t), history_fn, summary_fn)


def main():
    args = get_args()
    if args.test_mode:
        args.epochs, args.read_counts, args.window_size = 2, 1, 5
    try:
        start_time = time.time()
        shutil.rmtree(args.temp_dir, ignore_errors=True)
        for d in (args.temp_dir, args.output_dir, args.train_input_dir):
            check_path(d)
        train_preprocessing(args)
        check_path(args.train_model_dir)
        from nanoreviser_b200 import trainprep
        try:
            x_train, signal_x_train, y_train, y_train2 = trainprep.get_trainning_input(args.test_mode, args.train_input_dir, args.window_size)
        except Exception as e:
            raise RuntimeError(e)
        for tag, n_class, y, transfer in (('model1', 6, y_train, args.model1_train_dir), ('model2', 5, y_train2, args.model2_train_dir)):
            if args.model_type not in ('both', tag):
                continue
            try:
                if not args.test_mode:
                    print('[p:::] start to training %s, please waiting......' % tag)
                fit_and_save(args, tag, n_class, signal_x_train, x_train, y, transfer)
                if not args.test_mode:
                    print('[p:::] %s completed......' % tag.replace('model', 'model '))
            except Exception as e:
                raise RuntimeError('！！！[Error]training %s...... ' % tag.replace('model', 'model '), e)
        if not args.test_mode:
            print('[s:::] The training time of NanoReviser_train is :%.2f seconds' % (time.time() - start_time))
        else:
            for d in (args.output_dir, args.model_dir):
                shutil.rmtree(d, ignore_errors=True)
        shutil.rmtree(args.temp_dir, ignore_errors=True)
    except Exception as e:
        print(e)
        return 1
    return 0


if __name__ == '__main__':
    sys.exit(main())
