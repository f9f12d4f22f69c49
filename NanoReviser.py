#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""NanoReviser command line on B200: same flags as the reference CLI (NanoReviser.py:42-95 there),
same model files (./model/<S>/<S>_win13_50ep_model{1,2}.h5), same output files
(<o><stem>_out.fasta|fastq), but the revision path runs in libnrv.so (CUDA, sm_100a): no Keras,
no Guppy shell-out, no CPU fallback.

Differences that are deliberate:
  * the NN path is what runs (the shipped reference CLI builds un-weighted models and shells out to a
    missing Guppy binary -- SURVEY.md F1);
  * every fast5 in -d is processed (the reference silently drops ``len(files) % pool_size`` files,
    NanoReviser.py:212-219);
  * ``-e`` (failed reads file) is actually written;
  * extra flags ``--devices`` (comma separated GPU ids, one worker process per GPU) and
    ``--batch-bases`` (ragged batch budget per launch).
"""
import os
import shutil
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from optparse import OptionParser

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def get_args(argv=None):
    optParser = OptionParser(usage="%prog [-d] [-o]", version="%prog 1.0",
                             description="An Error-correction Tool for Nanopore Sequencing Based on a Deep Learning "
                                         "Algorithm (B200-native revision path)")
    optParser.add_option('-d', '--fast5_base_dir', action='store', type="string", dest='fast5_base_dir',
                         help='path to the fast5 files')
    optParser.add_option('-o', '--output_dir', action='store', type="string", dest='output_dir',
                         default='./unitest/nanorev_output/', help='path to store the output files')
    optParser.add_option('-F', '--output_format', action='store', type="string", dest='output_format',
                         default='fasta', help='format of the output files, default is fasta')
    optParser.add_option('-S', '--species', action='store', type="string", dest='species', default='human',
                         help='species of the model (ecoli or human), default is human')
    optParser.add_option("--thread", action="store", type="int", dest="thread", default=100,
                         help='host threads for fast5 ingest, default is 100 (capped at the core count)')
    optParser.add_option('-t', '--tmp_dir', action='store', type="string", dest='temp_dir', default='./unitest/tmp/',
                         help='path to the tmp dir (accepted for compatibility; nothing is staged there)')
    optParser.add_option('-e', '--failed_read', action='store', type="string", dest='failed_reads_filename',
                         default='failed_reads.txt', help='document to log the failed reads')
    optParser.add_option('-g', '--basecall_group', action='store', type="string", dest='basecall_group',
                         default='Basecall_1D_000', help='group holding the events table')
    optParser.add_option('-s', '--basecall_subgroup', action='store', type="string", dest='basecall_subgroup',
                         default='BaseCalled_template', help='subgroup holding the events table')
    optParser.add_option('--test_mode', action='store_true', default=False, help='just for unitest')
    optParser.add_option('--model1_predict_dir', action='store', type="string", dest='model1_predict_dir',
                         default='./model/human/human_win13_50ep_model1.h5', help='model dirs for model1')
    optParser.add_option('--model2_predict_dir', action='store', type="string", dest='model2_predict_dir',
                         default='./model/human/human_win13_50ep_model2.h5', help='model dirs for model2')
    optParser.add_option("-v", "--virsion", action="store_true", dest="virsion", help="version of NanoReviser")
    optParser.add_option('--devices', action='store', type="string", dest='devices', default='0',
                         help='CUDA device ids, e.g. 0,1,2 or 0-7; reads are sharded over them (one worker process per GPU, no collectives)')
    optParser.add_option('--batch-bases', action='store', type="int", dest='batch_bases', default=2_000_000,
                         help='ragged batch budget (bases per GPU launch)')
    optParser.add_option('--ingest', action='store', type="string", dest='ingest', default='native',
                         help='fast5 reader: native (C++ threads, default) or python')
    optParser.add_option('--slab-files', action='store', type="int", dest='slab_files', default=512,
                         help='fast5 files per ingest slab (the ingest thread works one slab ahead of the GPU)')
    optParser.add_option('--quiet', action='store_true', dest='quiet', default=False,
                         help='do not print one line per saved read')
    (tmp_args, _) = optParser.parse_args(argv)
    if tmp_args.virsion:
        print("The virsion of NanoReviser : 1.0 ")
        sys.exit()
    elif tmp_args.fast5_base_dir and tmp_args.output_dir:
        return tmp_args
    else:
        optParser.parse_args(['-h'])
        sys.exit()


def _test_logger():
    import logging
    logger = logging.getLogger('unitest')
    if not logger.handlers:
        logger.setLevel(logging.DEBUG)
        os.makedirs('./unitest', exist_ok=True)
        handler = logging.FileHandler('./unitest/unitest_log.txt', encoding='UTF-8')
        handler.setLevel(logging.INFO)
        handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
        logger.addHandler(handler)
        logger.addHandler(logging.StreamHandler())
    return logger


def _ingest(args, fn_sg):
    from nanoreviser_b200 import fast5
    path = os.path.join(args.fast5_base_dir, fn_sg)
    try:
        return fn_sg, fast5.read_fast5_arrays(path, args.basecall_group, args.basecall_subgroup), None
    except Exception as e:          # per-read try/except, like provide_fasta (NanoReviser.py:114-118)
        return fn_sg, None, e


def _write_read(args, fn_sg, bases, qul=None):
    from nanoreviser_b200 import api
    if not os.path.exists(args.output_dir):
        os.makedirs(args.output_dir)
    fast5_fn = os.path.join(args.fast5_base_dir, fn_sg)
    if args.output_format == 'fasta':
        api.prep_read_fasta(fast5_fn, api.out_filename(args.output_dir, fn_sg, 'fasta'), bases)
    else:
        api.prep_read_fastq(fast5_fn, api.out_filename(args.output_dir, fn_sg, 'fastq'), bases, qul)


def run_worker(args, file_list, device, logger=None):
    """One GPU, one pipeline (the reference fans single reads out to a multiprocessing.Pool, NanoReviser.py:203-223):

        ingest thread   native multi-threaded reader (C++, include/nrv.h nrv_ingest_fast5) slab by slab; files it does not
                        cover or cannot read go through the Python reader, which follows the reference branch by branch and
                        produces its error messages                                             -> bounded queue
        this thread     ragged batches by base budget (workqueue.make_batches), two batches in flight on the GPU
                        (Reviser.submit / wait: copies of one batch run under the kernels of the other)
        writer threads  per-read output files (prep_read_fasta / prep_read_fastq), fallback to the un-revised read on a
                        per-read failure (NanoReviser.py:146-154)."""
    import queue
    import threading
    t_worker0 = time.perf_counter()
    from nanoreviser_b200 import api, engine, fast5, weights, workqueue
    m1 = weights.load_model_weights(args.model1_predict_dir)
    m2 = weights.load_model_weights(args.model2_predict_dir)
    counts = {'ok': 0, 'fallback': 0, 'failed': 0}
    failed = []
    lock = threading.Lock()
    nthreads = max(1, min(int(args.thread), os.cpu_count() or 4, 32))
    native = getattr(args, 'ingest', 'native') == 'native'
    slab_files = int(getattr(args, 'slab_files', 512))
    fastq = args.output_format == 'fastq'

    def emit(fn_sg, ok, seq, orig_bases, qul=None):
        try:
            if ok:
                # D6': qualities of the two-model path come from the softmax outputs (include/nrv.h nrv_result.revised_qual)
                _write_read(args, fn_sg, seq, qul if fastq else None)
                with lock:
                    counts['ok'] += 1
            else:
                # fallback to the un-revised read (NanoReviser.py:146-154 / :172-181)
                if args.output_format == 'fasta':
                    _write_read(args, fn_sg, orig_bases.tobytes().decode('ascii'))
                else:
                    seq, qul = fast5.extract_fastq(os.path.join(args.fast5_base_dir, fn_sg), None)
                    _write_read(args, fn_sg, seq, qul)
                with lock:
                    counts['fallback'] += 1
                    failed.append(fn_sg)
            if logger:
                logger.info("Congratulations, NanoReviser is installed properly")
            elif not args.test_mode and not getattr(args, 'quiet', False):
                print('[p:::] ' + fn_sg.split('.')[0] + '_out.' + args.output_format + ' was saved......')
        except Exception as e:
            print('[！！！Error] stroring : ' + fn_sg.split('.')[0] + ' ' + str(e))
            with lock:
                failed.append(fn_sg)
            if logger:
                logger.error('[!!! Error] Basecalling')

    def load_slab(slab):
        """-> list of (Batch, names): the natively read files of the slab, then the ones the Python reader had to take"""
        units = []
        todo_python = list(slab)
        if native:
            paths = [os.path.join(args.fast5_base_dir, f) for f in slab]
            batch, fstatus, read_file, _a0, members = engine.ingest_fast5(paths, args.basecall_group, args.basecall_subgroup, nthreads,
                                                                          with_names=True)
            todo_python = [f for f, st in zip(slab, fstatus) if st != engine.INGEST_OK]
            if fastq and batch.n_reads and batch.qual is None:
                # the native reader attaches the basecaller's Phred scores when EVERY read of the slab has a Fastq dataset that
                # lines up; otherwise they are fetched per file here (bases that pass through unrevised keep them; reads where
                # the dataset is missing use Phred 40)
                def phred(i):
                    b0, b1 = int(batch.base_off[i]), int(batch.base_off[i + 1])
                    q = fast5.basecall_phred(paths[int(read_file[i])], batch.bases[b0:b1], args.basecall_group,
                                             args.basecall_subgroup)
                    return q if q is not None else np.full(b1 - b0, 40, np.uint8)
                with ThreadPoolExecutor(max_workers=nthreads) as ex:
                    batch.qual = np.concatenate(list(ex.map(phred, range(batch.n_reads))))
            if batch.n_reads:
                # a read out of a multi-read container is named after its member (read_<id>), any other after its file (:137)
                units.append((batch, [members[k] + '.fast5' if members[k] else slab[int(i)] for k, i in enumerate(read_file)]))
        if todo_python:
            with ThreadPoolExecutor(max_workers=nthreads) as ex:
                loaded = list(ex.map(lambda f: _ingest(args, f), todo_python))
            good = []
            for fn_sg, r, err in loaded:
                if r is None:
                    print('！！！[Error] fast5 file: ' + fn_sg.split('.')[0] + str(err))
                    with lock:
                        failed.append(fn_sg)
                        counts['failed'] += 1
                    if logger:
                        logger.error('[!!! Error] Basecalling')
                else:
                    if fastq:
                        q = fast5.basecall_phred(os.path.join(args.fast5_base_dir, fn_sg), r.bases, args.basecall_group,
                                                 args.basecall_subgroup)
                        r.qual = q if q is not None else np.full(r.n_bases, 40, np.uint8)
                    good.append((fn_sg, r))
            if good:
                units.append((engine.pack_batch([r for _, r in good]), [f for f, _ in good]))
        return units

    q_in = queue.Queue(maxsize=2)

    def ingest_loop():
        try:
            for s0 in range(0, len(file_list), slab_files):
                q_in.put(load_slab(file_list[s0:s0 + slab_files]))
        except BaseException as e:                     # surfaces in the GPU thread
            q_in.put(e)
        q_in.put(None)

    def write_batch(names, sub, out):
        for k, fn_sg in enumerate(names):
            ok = out.status[k] in (engine.NRV_READ_OK, engine.NRV_READ_TOO_SHORT)
            emit(fn_sg, ok, out.sequence(k), sub.bases[sub.base_off[k]:sub.base_off[k + 1]], out.quality(k) if fastq else None)

    t_ingest = threading.Thread(target=ingest_loop, name='nrv-ingest', daemon=True)
    t_ingest.start()
    writers = ThreadPoolExecutor(max_workers=max(2, min(8, nthreads)))
    pending_writes = []
    # where this thread's wall time goes (seconds): waiting for the ingest thread, for the GPU, for the writers
    stats = {'t_worker_start': t_worker0, 'wait_ingest_s': 0.0, 'wait_gpu_s': 0.0, 'wait_writers_s': 0.0, 'batches': 0, 'bases': 0}
    with engine.Reviser(m1, m2, device=device) as rv:
        stats['setup_s'] = time.perf_counter() - t_worker0        # weights, library, handle (CUDA context, weight re-packing)
        in_flight = []

        def drain_one():
            pend, names, sub = in_flight.pop(0)
            t0 = time.perf_counter()
            out = rv.wait(pend)
            t1 = time.perf_counter()
            pending_writes.append(writers.submit(write_batch, names, sub, out))
            while len(pending_writes) > 8:             # bound the memory held by finished batches
                pending_writes.pop(0).result()
            stats['wait_gpu_s'] += t1 - t0
            stats['wait_writers_s'] += time.perf_counter() - t1

        while True:
            t0 = time.perf_counter()
            units = q_in.get()
            stats['wait_ingest_s'] += time.perf_counter() - t0
            if 'first_slab_s' not in stats:
                stats['first_slab_s'] = time.perf_counter() - t_worker0
            if units is None:
                break
            if isinstance(units, BaseException):
                raise units
            for batch, names in units:
                lengths = np.diff(batch.base_off).tolist()
                for idx in workqueue.make_batches(range(batch.n_reads), lengths, int(args.batch_bases)):
                    # make_batches keeps the order: a batch is a contiguous range of the slab -> views, no copies
                    sub = batch if len(idx) == batch.n_reads else engine.slice_batch(batch, idx[0], idx[-1] + 1)
                    if len(in_flight) == 2:
                        drain_one()
                    in_flight.append((rv.submit(sub, want_qual=fastq), [names[i] for i in idx], sub))
                    stats['batches'] += 1
                    stats['bases'] += int(sub.base_off[-1])
        while in_flight:
            drain_one()
        stats['gpu_done_s'] = time.perf_counter() - t_worker0
    t0 = time.perf_counter()
    for f in pending_writes:
        f.result()
    writers.shutdown(wait=True)
    stats['wait_writers_s'] += time.perf_counter() - t0
    t_ingest.join()
    stats['total_s'] = time.perf_counter() - t_worker0
    return counts['ok'], counts['fallback'], counts['failed'], failed, stats


def _worker_entry(payload):
    args, files, device = payload
    # one process per GPU: expose only that GPU to the process before the CUDA driver initialises (cuInit enumerates every visible
    # device; with 8 B200s and 8 processes starting at once that alone took ~10 s per worker), then address it as device 0
    visible = os.environ.get('CUDA_VISIBLE_DEVICES')
    if visible:
        ids = [v for v in visible.split(',') if v != '']
        os.environ['CUDA_VISIBLE_DEVICES'] = ids[device] if device < len(ids) else str(device)
    else:
        os.environ['CUDA_VISIBLE_DEVICES'] = str(device)
    return run_worker(args, files, 0, _test_logger() if args.test_mode else None)


def main(ar_args):
    logger = None
    if ar_args.test_mode:
        logger = _test_logger()
        ar_args.model1_predict_dir = './model/ecoli/ecoli_win13_50ep_model1.h5'
        ar_args.model2_predict_dir = './model/ecoli/ecoli_win13_50ep_model2.h5'
    if ar_args.species:   # NanoReviser.py:191-193: -S (default 'human') overrides the model paths, also in test mode
        ar_args.model1_predict_dir = './model/' + str(ar_args.species) + '/' + str(ar_args.species) + '_win13_50ep_model1.h5'
        ar_args.model2_predict_dir = './model/' + str(ar_args.species) + '/' + str(ar_args.species) + '_win13_50ep_model2.h5'
    if not (os.path.exists(ar_args.model1_predict_dir) and os.path.exists(ar_args.model2_predict_dir)):
        raise RuntimeError('！！！[Error] model file: Please check the dir of models file!!')
    os.makedirs(ar_args.output_dir, exist_ok=True)
    fast5_fns = [f for f in sorted(os.listdir(ar_args.fast5_base_dir)) if not f.startswith('.')]
    devices = []
    for d in str(ar_args.devices).split(','):          # "0,1,2" or "0-7"
        if '-' in d:
            a, b = d.split('-')
            devices += list(range(int(a), int(b) + 1))
        elif d != '':
            devices.append(int(d))
    start_time = time.time()
    if len(devices) <= 1:
        res = [run_worker(ar_args, fast5_fns, devices[0] if devices else 0, logger)]
    else:
        # reads shard naturally: length-balanced partition by file size (proxy for bases), one process per GPU
        import multiprocessing as mp
        from nanoreviser_b200 import workqueue
        sizes = [os.path.getsize(os.path.join(ar_args.fast5_base_dir, f)) for f in fast5_fns]
        parts = workqueue.lpt_partition(sizes, len(devices))
        ctx = mp.get_context('spawn')
        with ctx.Pool(len(devices)) as pool:
            res = pool.map(_worker_entry, [(ar_args, [fast5_fns[i] for i in p], d) for p, d in zip(parts, devices)])
    failed = [f for r in res for f in r[3]]
    if failed and not ar_args.test_mode:
        with open(os.path.join(ar_args.output_dir, ar_args.failed_reads_filename), 'w') as fp:
            fp.write('\n'.join(failed) + '\n')
    end_time = time.time()
    if not ar_args.test_mode:
        print('[s:::] All reads done: %d revised, %d written un-revised, %d unreadable.' % (
            sum(r[0] for r in res), sum(r[1] for r in res), sum(r[2] for r in res)))
        print('[s:::] NanoReviser time consuming:%.2f seconds' % (end_time - start_time))
    else:
        shutil.rmtree(ar_args.output_dir, ignore_errors=True)    # NanoReviser.py:231-232
    return res


if __name__ == '__main__':
    ar_args = get_args()
    try:
        main(ar_args)
    except Exception as e:
        print(e)
