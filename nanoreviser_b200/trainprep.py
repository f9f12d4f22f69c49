"""Training-label generation (host side): from a basecalled read and its alignment to the genome to
the per-base training arrays and the windowed training tensors.

Mirrors, function by function (same names, arguments, return order and error behaviour):

  parse_fasta                        nanorevutils/input_handeler.py:28-57
  comp_base / rev_comp               nanorevutils/alignutils.py:63-74
  load_sam_record                    nanorevutils/alignutils.py:44-61 (the part of align_to_genome after the mapper has run)
  align_to_genome                    nanorevutils/alignutils.py:31-61 (runs the external mapper, graphmap by default)
  parse_sam_record                   nanorevutils/alignutils.py:77-177
  fix_raw_starts_for_clipped_bases   nanorevutils/preprocessing.py:18-42
  clean_read_map_ref                 nanorevutils/preprocessing.py:45-82
  get_base_color / get_base_label    nanorevutils/preprocessing.py:173-180
  read_training_arrays               nanorevutils/nanorevtrainutils.py:56-130 (handel_input_fast5 without the file handling); the
                                     re-segmentation (signal_segmentation) runs on the GPU through the same K1 kernels as inference
  get_trainning_input                nanorevutils/nanorevtrainutils.py:146-218

The reference walks Python lists character by character; here the alignment columns are numpy byte arrays and the cleaning /
windowing steps are vectorised.  Pinned against the reference's own functions on randomised alignments (both strands, H / S
clips, I / D / N / P / = / X operations, leading and trailing non-match operations): tests/golden/make_trainprep_fixtures.py ->
tests/golden/trainprep.json, tests/test_trainprep.py.  Two quirks of the reference are kept on purpose and tested: a trailing
non-match operation adds the length of the FIRST remaining operation to end_clipped_bases (alignutils.py:137), and the last
alignment column is always kept by clean_read_map_ref, also when it is a deletion (preprocessing.py:75-78).
"""
from __future__ import annotations

import os
import re
from subprocess import call
from typing import Dict, List, Tuple

import numpy as np

SAM_FIELDS = ('qName', 'flag', 'rName', 'pos', 'mapq', 'cigar', 'rNext', 'pNext', 'tLen', 'seq', 'qual')
_CIGAR_PAT = re.compile(r'(\d+)([MIDNSHP=X])')
_COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A', '-': '-', 'N': 'N'}
_COLOR = {'A': 250, 'G': 180, 'T': 100, 'C': 30}
_LABEL = {'A': 5, 'G': 4, 'T': 3, 'C': 2, '-': 1, 'D': 0}


def get_base_color(base) -> int:
    return _COLOR.get(base, 0)


def get_base_label(base) -> int:
    return _LABEL.get(base, 0)


def _lut(table: Dict[str, int]) -> np.ndarray:
    lut = np.zeros(256, np.int64)
    for k, v in table.items():
        lut[ord(k)] = v
    return lut


_COLOR_LUT, _LABEL_LUT = _lut(_COLOR), _lut(_LABEL)


def base_colors(chars: np.ndarray) -> np.ndarray:
    """get_base_color over a uint8 array of characters"""
    return _COLOR_LUT[np.asarray(chars, np.uint8)]


def base_labels(chars: np.ndarray) -> np.ndarray:
    """get_base_label over a uint8 array of characters"""
    return _LABEL_LUT[np.asarray(chars, np.uint8)]


def parse_fasta(fasta_fn: str) -> Dict[str, str]:
    records: Dict[str, str] = {}
    curr_id, chunks = None, []
    with open(fasta_fn, 'r') as fp:
        for line in fp:
            if line.startswith('>'):
                if curr_id is not None and chunks:
                    records[curr_id] = ''.join(chunks)
                chunks = []
                curr_id = line.replace('>', '').strip().split()[0]
            else:
                s = line.strip()
                if s:
                    chunks.append(s)
    if curr_id is not None and chunks:
        records[curr_id] = ''.join(chunks)
    return records


def comp_base(base: str) -> str:
    return _COMP.get(base, 'N')


def rev_comp(seq: str) -> str:
    return ''.join(comp_base(b) for b in seq[::-1])


def prep_graphmap_options(genome_fn, read_fn, out_fn, output_format, num_align_ps):
    return ['align', '-r', genome_fn, '-d', read_fn, '-o', out_fn, '-t', str(num_align_ps)]


def load_sam_record(align_output: List[str]) -> Dict[str, str]:
    """The LAST alignment line of a SAM file as a field dictionary; the reference's two error cases."""
    rec: Dict[str, str] = {}
    for line in align_output:
        if line.startswith('@'):
            continue
        rec = dict(zip(SAM_FIELDS, line.strip().split()))
    if not rec:
        raise RuntimeError('Map Error, there is no read record in the sam file')
    if len(rec) < len(SAM_FIELDS) or rec['rName'] == '*':
        raise RuntimeError('Map Error, the read is unmapped.')
    return rec


def align_to_genome(out_fn, graphmap_exe, mapper_options):
    with open(os.devnull, 'w') as sink:
        exit_status = call([graphmap_exe] + list(mapper_options), stdout=sink, stderr=sink)
    if exit_status != 0:
        raise RuntimeError('Align Error, please check your graphmap or bwa mem')
    with open(out_fn, 'r') as fp:
        return load_sam_record(fp.readlines())


def parse_sam_record(r_sam_record, genome_index):
    """-> (readVals, refVals, mapVals, genomeLoc, start_clipped_bases, end_clipped_bases); the three columns are lists of
    one-character strings like the reference's (use ``alignment_columns`` for the byte arrays)."""
    rv, fv, mv, loc, sc, ec = alignment_columns(r_sam_record, genome_index)
    as_list = lambda a: list(a.tobytes().decode('ascii'))
    return as_list(rv), as_list(fv), as_list(mv), loc, sc, ec


def alignment_columns(r_sam_record, genome_index) -> Tuple[np.ndarray, np.ndarray, np.ndarray, dict, int, int]:
    cigar = [(int(n), t) for n, t in _CIGAR_PAT.findall(r_sam_record['cigar'])]
    if len(cigar) < 1:
        raise RuntimeError('Invalid cigar string produced.')
    strand = '-' if int(r_sam_record['flag']) & 0x10 else '+'
    if strand == '-':
        cigar = cigar[::-1]
    q_seq = str(r_sam_record['seq'] if strand == '+' else rev_comp(r_sam_record['seq']))
    start_clipped = end_clipped = 0
    if cigar[0][1] == 'H':
        start_clipped += cigar[0][0]; cigar = cigar[1:]
    if cigar[-1][1] == 'H':
        end_clipped += cigar[-1][0]; cigar = cigar[:-1]
    if cigar[0][1] == 'S':
        start_clipped += cigar[0][0]; q_seq = q_seq[cigar[0][0]:]; cigar = cigar[1:]
    if cigar[-1][1] == 'S':
        end_clipped += cigar[-1][0]; q_seq = q_seq[:-cigar[-1][0]]; cigar = cigar[:-1]
    t_len = sum(n for n, t in cigar if t in 'MDN=X')
    pos = int(r_sam_record['pos'])
    t_seq = str(genome_index[r_sam_record['rName']][pos - 1:pos + t_len - 1])
    if strand == '-':
        t_seq = rev_comp(t_seq)
    # the alignment has to start and end with matched bases
    while cigar[0][1] not in 'M=X':
        if cigar[0][1] in 'IP':
            t_seq = t_seq[cigar[0][0]:]
        else:
            q_seq = q_seq[cigar[0][0]:]
            start_clipped += cigar[0][0]
        cigar = cigar[1:]
    while cigar[-1][1] not in 'M=X':
        if cigar[-1][1] in 'IP':
            t_seq = t_seq[:-cigar[-1][0]]
        else:
            q_seq = q_seq[:-cigar[-1][0]]
            end_clipped += cigar[0][0]                   # (sic: alignutils.py:137 adds the FIRST operation's length)
        cigar = cigar[:-1]
    q_len = sum(n for n, t in cigar if t in 'MIP=X')
    assert len(q_seq) == q_len, 'Read sequence from SAM and cooresponding cigar string do not agree.'
    total = sum(n for n, _ in cigar)
    q = np.frombuffer(q_seq.encode('ascii'), np.uint8)
    t = np.frombuffer(t_seq.encode('ascii'), np.uint8)
    read_v = np.full(total, ord('-'), np.uint8)
    ref_v = np.full(total, ord('-'), np.uint8)
    map_v = np.empty(total, np.uint8)
    o = qi = ti = 0
    for n, typ in cigar:
        if typ in 'M=X':
            qs, ts = q[qi:qi + n], t[ti:ti + n]
            m = min(len(qs), len(ts))                    # the reference zips: a short reference slice shortens the block
            read_v[o:o + len(qs)] = qs
            if len(qs) != len(ts):                       # (cannot happen for a consistent record; keep the reference's zip semantics)
                raise AssertionError('reference slice shorter than the cigar block')
            ref_v[o:o + m] = ts[:m]
            map_v[o:o + m] = np.where(qs[:m] == ts[:m], ord('M'), ord('X'))
            qi += n; ti += n
        elif typ in 'IP':
            read_v[o:o + n] = q[qi:qi + n]
            map_v[o:o + n] = ord('I')
            qi += n
        else:
            ref_v[o:o + n] = t[ti:ti + n]
            map_v[o:o + n] = ord('D')
            ti += n
        o += n
    loc = {'Start': pos - 1, 'Strand': strand, 'Chrom': r_sam_record['rName']}
    return read_v, ref_v, map_v, loc, start_clipped, end_clipped


def fix_raw_starts_for_clipped_bases(start_clipped_bases, end_clipped_bases, starts_rel_to_read, event_length,
                                     read_start_rel_to_raw, ab_p_model_states, ab_weights):
    if start_clipped_bases > 0:
        k = int(start_clipped_bases)
        start_clipped_obs = int(starts_rel_to_read[k])
        ab_p_model_states = ab_p_model_states[k:]
        ab_weights = ab_weights[k:]
        event_length = event_length[k:]
        starts_rel_to_read = starts_rel_to_read[k:] - start_clipped_obs
        read_start_rel_to_raw = int(read_start_rel_to_raw + start_clipped_obs)
    if end_clipped_bases > 0:
        k = int(end_clipped_bases)
        starts_rel_to_read = starts_rel_to_read[:-k]
        ab_p_model_states = ab_p_model_states[:-k]
        ab_weights = ab_weights[:-k]
        event_length = event_length[:-k]
    return starts_rel_to_read, event_length, read_start_rel_to_raw, ab_p_model_states, ab_weights


_MXI = np.zeros(256, bool)
for _c in 'MXI':
    _MXI[ord(_c)] = True


def clean_columns(read_v: np.ndarray, map_v: np.ndarray, ref_v: np.ndarray):
    """clean_read_map_ref on byte arrays: every alignment column that carries a read base is kept; a column followed by a
    deletion gets map 'D' and label 'D' in refVals (model 1: 'a base is missing after this one') and keeps its own reference
    base in refVals2 (model 2); deletion columns are dropped -- except the last column, which the reference always keeps."""
    read_v, map_v, ref_v = (np.asarray(a, np.uint8) for a in (read_v, map_v, ref_v))
    m1, m2 = map_v[:-1], map_v[1:]
    keep = _MXI[m1] & (_MXI[m2] | (m2 == ord('D')))
    next_d = m2 == ord('D')
    out_read = np.concatenate([read_v[:-1][keep], read_v[-1:]])
    out_map = np.concatenate([np.where(next_d, ord('D'), m1)[keep], map_v[-1:]]).astype(np.uint8)
    out_ref = np.concatenate([np.where(next_d, ord('D'), ref_v[:-1])[keep], ref_v[-1:]]).astype(np.uint8)
    out_ref2 = np.concatenate([ref_v[:-1][keep], ref_v[-1:]])
    return out_read, out_map, out_ref, out_ref2


def clean_read_map_ref(readVals, mapVals, refVals):
    """-> (clean_readVals, clean_mapVals, clean_refVals, clean_refVals2) as lists of characters (reference signature)."""
    enc = lambda v: np.frombuffer(''.join(v).encode('ascii'), np.uint8)
    out = clean_columns(enc(readVals), enc(mapVals), enc(refVals))
    return tuple(list(a.tobytes().decode('ascii')) for a in out)


def read_training_arrays(read, sam_record, genome_index, reviser) -> Dict[str, np.ndarray]:
    """The arrays handel_input_fast5 saves per read (nanorevtrainutils.py:104-117).  `read` = fast5.ReadArrays (get_read_data),
    `reviser` = engine.Reviser: the re-segmentation (per-base mean / std, normalised 50-sample windows, shift, scale) runs on the
    GPU through nrv_segment, bit-identical to preprocessing.signal_segmentation (tests/test_gpu_parity.py)."""
    from . import engine
    read_v, ref_v, map_v, _loc, sc, ec = alignment_columns(sam_record, genome_index)
    length = np.asarray(read.length)
    starts, length, a0, ev_mean, ev_std = fix_raw_starts_for_clipped_bases(
        int(sc), int(ec), np.asarray(read.starts), length, int(read.a0), np.asarray(read.ev_mean), np.asarray(read.ev_std))
    c_read, c_map, c_ref, c_ref2 = clean_columns(read_v, map_v, ref_v)
    n = len(starts)
    b = engine.Batch(signal=np.ascontiguousarray(read.signal[int(a0):], np.int16),
                     sig_off=np.array([0, len(read.signal) - int(a0)], np.int64),
                     starts=np.ascontiguousarray(starts, np.int32), base_off=np.array([0, n], np.int64),
                     bases=np.ascontiguousarray(np.asarray(read.bases)[int(sc):int(sc) + n], np.uint8),
                     ev_mean=np.ascontiguousarray(ev_mean, np.float32), ev_std=np.ascontiguousarray(ev_std, np.float32),
                     last_dur=np.array([int(length[-1])], np.int32))
    shift, scale, seg_mean, seg_std, _x, win, _status = reviser.segment(b, want_windows=True)
    return dict(refvals=base_labels(c_ref), refvals2=base_labels(c_ref2), readVals=base_colors(c_read),
                signal_mean=seg_mean, signal_std=seg_std, signal_len=np.asarray(length), ab_mean=np.asarray(ev_mean),
                ab_std=np.asarray(ev_std), signal_x=win.astype(np.float64), mapvals=c_map, starts=np.asarray(starts),
                scale=np.float64(scale[0]), shift=np.float64(shift[0]))


def training_tensors(arrays: List[Dict[str, np.ndarray]], window_size: int = 13):
    """get_trainning_input from in-memory per-read dictionaries (the .npz contents): reads are concatenated base by base, then
    every run of `window_size` consecutive bases is one sample labelled by its centre base (nanorevtrainutils.py:160-214)."""
    xs, sigs, ys, y2s = [], [], [], []
    for a in arrays:
        cols = np.vstack([np.asarray(a['readVals']) / 300.0, np.asarray(a['signal_mean']) / a['shift'],
                          np.asarray(a['signal_std']) / a['scale'], np.asarray(a['signal_len']) / 10.0,
                          np.asarray(a['ab_mean']), np.asarray(a['ab_std'])])
        xs.append(cols); sigs.append(np.asarray(a['signal_x'])); ys.append(np.asarray(a['refvals'])); y2s.append(np.asarray(a['refvals2']))
    x = np.concatenate(xs, axis=-1).T.astype(float)
    signal_x = np.concatenate(sigs, axis=0)
    y = np.concatenate(ys).astype(float)
    y2 = np.concatenate(y2s).astype(float)
    W = int(window_size)
    assert len(x) > 2 * W
    idx = np.arange(len(x) - W)[:, None] + np.arange(W)[None, :]
    x_train = x[idx]
    sidx = np.arange(len(signal_x) - W)[:, None] + np.arange(W)[None, :]
    signal_x_train = signal_x[sidx].astype(float)
    bef, aft = (W - 1) // 2, (W + 1) // 2
    y_train = y[bef:-aft].reshape(-1, 1)
    y_train2 = (y2[bef:-aft] - 1).reshape(-1, 1)
    return x_train, signal_x_train, y_train, y_train2


def get_trainning_input(test_mode, train_input_dir, window_size=13):
    """The reference entry point: every readable .npz of the directory, in os.listdir order."""
    arrays = []
    for fn in os.listdir(train_input_dir):
        path = os.path.join(train_input_dir, fn)
        if not path.endswith('.npz'):
            continue
        try:
            z = np.load(path)
            arrays.append({k: z[k] for k in ('shift', 'scale', 'readVals', 'signal_mean', 'signal_std', 'signal_len', 'ab_mean',
                                             'ab_std', 'signal_x', 'refvals', 'refvals2')})
        except Exception:
            print('！！！[Error] training input file:', path)
            continue
    try:
        out = training_tensors(arrays, window_size)
    except AssertionError:
        raise
    except Exception:
        raise RuntimeError('！！！[Error] fatal errors in loading training data.')
    if not test_mode:
        print('[p:::] input files has been load......')
    return out
