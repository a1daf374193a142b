"""Per-read detail output, its index files and ``--predDet 0``.

What the reference writes for every predicted read (``bin/DeepMod_scripts/myDetect.py``):

* ``:716-753``  one HDF5 group ``/pred/pred_<i>`` per read in ``<ctfolder>/rnn.pred.detail.fast5.<batchid>``: attributes
  (mapped_chr, mapped_strand, mapped_start, mapped_end, clipped_bases_start/end, num_insertions, num_deletions,
  num_matches, num_mismatches, pred_mod_num, f5file, readk) and the gzip'd compound dataset ``predetail`` =
  ``base_map_info`` with dtype ``[refbase S1, readbase S1, refbasei u8, readbasei u8, mod_pred int]``;
* ``:762-782``  per batch and chromosome a text index ``<ctfolder>/<chr>.rnn.pred.ind.<batchid>``, one sorted line per read:
  ``chr strand pos pred_key fast5-relative-path detail-file-relative-path`` (single spaces, trailing space);
* ``:1193-1221`` the merged ``<outFolder><FileID>/rnn.pred.ind.<chr>`` with two ``#base_folder_*`` header lines;
* ``:989-1023``, ``:1232-1263`` ``--predDet 0 --predpath``: ``read_file_list`` + ``read_pred_detail`` feed ``sum_handler``.

This image has no HDF5 library (no h5py, no libhdf5), so an HDF5 writer could not be validated against a real reader.
The detail RECORDS and ATTRIBUTES are therefore stored, field for field, in a documented container -- one
``rnn.pred.detail.dmpd.<batchid>`` per batch: a zip of ``.npy`` members (``numpy.savez_compressed``), see ``CONTAINER``
below -- the index files are byte-for-byte the reference's, and ``to_hdf5`` / ``read_detail`` convert to and from the
reference's HDF5 layout wherever h5py exists.  Known deviations, all in fields the summary never reads: ``readbasei``
counts from the first kept base (the packed input does not carry the read offset of the first match), the indel /
mismatch attributes count the stored (trimmed) columns, and index positions are the alignment start after clip removal
(equal to SAM POS-1 unless a CIGAR starts with D/N/X).
"""
import glob
import os
import zipfile
from collections import defaultdict

import numpy as np

from . import capi, checkpoint

PRE_BASE_STR = "rnn.pred.ind"                                   # myDetect.py:40
DETAIL_NAME = "rnn.pred.detail.dmpd"
DETAIL_DTYPE = np.dtype([("refbase", "S1"), ("readbase", "S1"), ("refbasei", "<u8"), ("readbasei", "<u8"), ("mod_pred", "<i8")])
CONTAINER = """rnn.pred.detail.dmpd.<batchid>: numpy .npz (zip, deflate) with, for the n reads of the batch,
  keys[n] 'pred_<i>' | rec_off[n+1] | refbase, readbase u1[rec_off[n]] | refbasei, readbasei u8 | mod_pred i1
  mapped_chr[n], mapped_strand[n] ('+'/'-'), mapped_start, mapped_end, clipped_bases_start, clipped_bases_end,
  num_insertions, num_deletions, num_matches, num_mismatches, pred_mod_num [n] | f5file[n], readk[n] | contig_len[n]"""
_ATTR_INT = ("mapped_start", "mapped_end", "clipped_bases_start", "clipped_bases_end", "num_insertions", "num_deletions",
             "num_matches", "num_mismatches", "pred_mod_num")


def column_predictions(pb, pred, status):
    """mPredict1's label write-back (:822-833) for a whole packed batch: prediction k of a read belongs to its k-th
    non-gap alignment column; gaps and rejected reads get 0.  -> int8 [n_cols]"""
    readb = pb.a["col_readbase"]
    col_off = pb.a["col_off"]
    nongap = readb != ord("-")
    rank = np.cumsum(nongap) - nongap                                    # exclusive rank over the batch
    win_off = np.concatenate([[0], np.cumsum(pb.n_windows_per_read)])
    n_cols_per_read = np.diff(col_off)
    read_of_col = np.repeat(np.arange(pb.n_reads), n_cols_per_read)
    k = rank - np.repeat(rank[np.minimum(col_off[:-1], max(len(rank) - 1, 0))] if len(rank) else rank[:0], n_cols_per_read)
    ok = nongap & (np.repeat(status == capi.READ_OK, n_cols_per_read)) & (k < np.repeat(pb.n_windows_per_read, n_cols_per_read))
    out = np.zeros(len(readb), np.int8)
    w = (win_off[read_of_col] + k)[ok]
    out[ok] = pred[w]
    return out


class DetailWriter(object):
    """Writes the per-read detail containers and per-batch index files of one rank (its ``ctfolder`` = the rank)."""

    def __init__(self, out_dir, wrk_base, rank=0, contig_len=None):
        self.out_dir, self.wrk_base, self.contig_len = out_dir, wrk_base, contig_len
        self.ctfolder = os.path.join(out_dir, str(rank))
        os.makedirs(self.ctfolder, exist_ok=True)
        self.batchid = 0

    def add_batch(self, path, first, pb, pred, status, contig_names):
        contig_len = self.contig_len
        ok = np.flatnonzero(status == capi.READ_OK)
        if len(ok) == 0:
            return None
        mod = column_predictions(pb, pred, status)
        col_off = pb.a["col_off"]
        refb, readb, refpos = pb.a["col_refbase"], pb.a["col_readbase"], pb.a["col_refpos"]
        seg = [slice(int(col_off[r]), int(col_off[r + 1])) for r in ok]
        lens = np.array([s.stop - s.start for s in seg], np.int64)
        cat = lambda a: np.concatenate([a[s] for s in seg])
        rb, qb, rp, mp = cat(refb), cat(readb), cat(refpos).astype(np.uint64), cat(mod)
        rec_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        nongap = qb != ord("-")
        rbi = (np.cumsum(nongap) - nongap)
        rbi = (rbi - np.repeat(rbi[rec_off[:-1]], lens)).astype(np.uint64)
        strand = pb.a["strand"][ok]
        fwd = strand >= 0
        first_pos, last_pos = rp[rec_off[:-1]], rp[rec_off[1:] - 1]
        seg_sum = lambda m: np.add.reduceat(m.astype(np.int64), rec_off[:-1])
        n_ins, n_del = seg_sum(rb == ord("-")), seg_sum(qb == ord("-"))
        n_match = seg_sum((rb == qb) & nongap)
        rel = os.path.relpath(path, self.wrk_base) if self.wrk_base else os.path.basename(path)
        ids = pb.extra.get("read_id")
        f5file = [("%s#%s" % (rel, ids[int(r)].decode() if ids is not None else int(first[int(r)]))) for r in ok]
        aln_pos = pb.extra.get("aln_pos")
        ind_pos = aln_pos[ok].astype(np.int64) if aln_pos is not None else np.where(fwd, first_pos, last_pos).astype(np.int64)
        chroms = [contig_names[int(c)] for c in pb.a["contig"][ok]]
        arrays = dict(
            keys=np.array(["pred_%d" % int(first[int(r)]) for r in ok]), rec_off=rec_off, refbase=rb, readbase=qb, refbasei=rp,
            readbasei=rbi, mod_pred=mp, mapped_chr=np.array(chroms), mapped_strand=np.where(fwd, "+", "-"),
            mapped_start=np.where(fwd, first_pos, last_pos), mapped_end=np.where(fwd, last_pos, first_pos),
            clipped_bases_start=pb.a["start_clip"][ok], clipped_bases_end=pb.a["end_clip"][ok], num_insertions=n_ins,
            num_deletions=n_del, num_matches=n_match, num_mismatches=lens - n_ins - n_del - n_match,
            pred_mod_num=seg_sum(mp == 1), f5file=np.array(f5file), readk=np.array(f5file),
            contig_len=np.array([int(contig_len[int(c)]) if contig_len is not None else 0 for c in pb.a["contig"][ok]], np.int64))
        name = "%s.%d" % (DETAIL_NAME, self.batchid)
        with open(os.path.join(self.ctfolder, name), "wb") as fh:
            np.savez_compressed(fh, **arrays)
        pred_rel = os.path.relpath(os.path.join(self.ctfolder, name), self.out_dir)
        # index lines: sorted list-of-lists exactly like sp_options['Mod'] (:719, :762)
        mod_list = sorted([chroms[i], "+" if fwd[i] else "-", int(ind_pos[i]), str(arrays["keys"][i]), f5file[i], pred_rel]
                          for i in range(len(ok)))
        cur_chr, writer = None, None
        for mfi in mod_list:
            if cur_chr != mfi[0]:
                if writer is not None:
                    writer.close()
                cur_chr = mfi[0]
                writer = open(os.path.join(self.ctfolder, "%s.%s.%d" % (cur_chr, PRE_BASE_STR, self.batchid)), "w")
            writer.write(" ".join([str(x) for x in mfi] + ["\n"]))
        if writer is not None:
            writer.close()
        self.batchid += 1
        return os.path.join(self.ctfolder, name)

    def close(self):
        pass


def merge_index_files(out_dir, wrk_base):
    """``<out_dir>/*/<chr>.rnn.pred.ind.<batch>`` -> ``<out_dir>/rnn.pred.ind.<chr>`` (myDetect.py:1193-1221)."""
    chr_dict = defaultdict(list)
    for pcf in glob.glob(os.path.join(out_dir, "*/*." + PRE_BASE_STR + ".*")):
        chr_dict[pcf.split("/")[-1].split("." + PRE_BASE_STR)[0]].append(pcf)
    written = []
    for ck in chr_dict:
        cur_list = [["#base_folder_fast5", wrk_base], ["#base_folder_output", os.path.abspath(out_dir)]]
        for sub in chr_dict[ck]:
            with open(sub) as fh:
                for line in fh:
                    line = line.strip()
                    if line:
                        lsp = line.split()
                        lsp[2] = int(lsp[2])
                        cur_list.append(lsp)
        cur_list = sorted(cur_list)
        path = os.path.join(out_dir, PRE_BASE_STR + "." + ck)
        with open(path, "w") as fh:
            for mfi in cur_list:
                fh.write(" ".join([str(x) for x in mfi] + ["\n"]))
        written.append(path)
    return sorted(written)


def read_file_list(cur_cif, cur_strand):
    """Restatement of ``read_file_list`` (:989-1010) -> (entries of that strand, base_folder_output)."""
    base_out, out = None, []
    with open(cur_cif) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            lsp = line.split()
            if line[0] == "#":
                if lsp[1][0] not in "/\\":                  # (sic: the reference tests the FIRST character, :1000)
                    lsp[1] = lsp[1] + "/"
                if lsp[0] == "#base_folder_output":
                    base_out = lsp[1]
            elif lsp[1] == cur_strand:
                out.append(lsp)
    return out, base_out


class _Detail(object):
    """One detail file: ``get(key) -> (refbase, readbase, refbasei, mod_pred, chr, strand)``; ``contig_len(keys)``."""

    def __init__(self, path):
        self.path, self.z = path, None
        if zipfile.is_zipfile(path):
            with np.load(path, allow_pickle=False) as f:
                self.z = {k: f[k] for k in f.files}
            self.index = {str(k): i for i, k in enumerate(self.z["keys"])}

    def get(self, key):
        if self.z is None:
            return read_detail_hdf5(self.path, key)[:6]          # a file the reference wrote
        z, i = self.z, self.index[key]
        a, b = int(z["rec_off"][i]), int(z["rec_off"][i + 1])
        return (z["refbase"][a:b], z["readbase"][a:b], z["refbasei"][a:b].astype(np.int64), z["mod_pred"][a:b].astype(np.int8),
                str(z["mapped_chr"][i]), str(z["mapped_strand"][i]))

    def contig_len(self, keys):
        """Length of the contig these reads map to: stored in the container; for the reference's files (which do not
        record it) the largest stored position + 1."""
        if self.z is not None:
            return max([int(self.z["contig_len"][self.index[k]]) for k in keys] + [0])
        return max([int(self.get(k)[2].max()) + 1 for k in keys] + [0])


def read_detail_hdf5(path, key):
    """``read_pred_detail`` (:1015-1026) on a file the REFERENCE wrote; needs h5py."""
    try:
        import h5py
    except ImportError:
        raise capi.DeepModError("%s is an HDF5 detail file of the reference; reading it needs h5py, which is not installed" % path)
    with h5py.File(path, "r") as mr:
        grp = mr["/pred/%s" % key]
        m = grp["predetail"][()]
        chrom, strand = grp.attrs["mapped_chr"], grp.attrs["mapped_strand"]
    dec = lambda x: x.decode() if isinstance(x, bytes) else str(x)
    return (np.frombuffer(m["refbase"].astype("S1").tobytes(), np.uint8), np.frombuffer(m["readbase"].astype("S1").tobytes(), np.uint8),
            m["refbasei"].astype(np.int64), m["mod_pred"].astype(np.int8), dec(chrom), dec(strand), 0)


def to_hdf5(detail_path, out_path):
    """Convert one ``rnn.pred.detail.dmpd.<batch>`` into the reference's ``rnn.pred.detail.fast5.<batch>`` (:722-753)."""
    try:
        import h5py
    except ImportError:
        raise capi.DeepModError("to_hdf5 needs h5py")
    with np.load(detail_path, allow_pickle=False) as z, h5py.File(out_path, "a") as out:
        base = out.require_group("pred")
        off = z["rec_off"]
        cols = {k: z[k] for k in ("refbase", "readbase", "refbasei", "readbasei", "mod_pred")}
        for i, key in enumerate(z["keys"]):
            key = str(key)
            if key in base:
                del base[key]
            g = base.create_group(key)
            g.attrs["mapped_chr"] = str(z["mapped_chr"][i])
            g.attrs["mapped_strand"] = str(z["mapped_strand"][i])
            for a in _ATTR_INT:
                g.attrs[a] = int(z[a][i])
            g.attrs["f5file"] = str(z["f5file"][i])
            g.attrs["readk"] = str(z["readk"][i])
            a, b = int(off[i]), int(off[i + 1])
            rec = np.zeros(b - a, DETAIL_DTYPE)
            rec["refbase"] = cols["refbase"][a:b].view("S1")
            rec["readbase"] = cols["readbase"][a:b].view("S1")
            rec["refbasei"], rec["readbasei"], rec["mod_pred"] = cols["refbasei"][a:b], cols["readbasei"][a:b], cols["mod_pred"][a:b]
            g.create_dataset("predetail", data=rec, compression="gzip")


def records_of(detail_path):
    """All reads of a container as the reference's structured arrays: {key: (attrs dict, predetail[DETAIL_DTYPE])}."""
    out = {}
    with np.load(detail_path, allow_pickle=False) as z:
        off = z["rec_off"]
        for i, key in enumerate(z["keys"]):
            a, b = int(off[i]), int(off[i + 1])
            rec = np.zeros(b - a, DETAIL_DTYPE)
            rec["refbase"], rec["readbase"] = z["refbase"][a:b].view("S1"), z["readbase"][a:b].view("S1")
            rec["refbasei"], rec["readbasei"], rec["mod_pred"] = z["refbasei"][a:b], z["readbasei"][a:b], z["mod_pred"][a:b]
            attrs = {k: int(z[k][i]) for k in _ATTR_INT}
            attrs.update(mapped_chr=str(z["mapped_chr"][i]), mapped_strand=str(z["mapped_strand"][i]), f5file=str(z["f5file"][i]),
                         readk=str(z["readk"][i]))
            out[str(key)] = (attrs, rec)
    return out


def summarise_stored(moptions):
    """``--predDet 0 --predpath <dir>``: the summary phase of mDetect_manager (:1232-1263) from stored per-read
    predictions; the per-position accumulation (:1089-1100) runs on the GPU (``dm_accumulate_records``)."""
    predpath = moptions["predpath"]
    out_dir = moptions["outFolder"] + moptions["FileID"]
    os.makedirs(out_dir, exist_ok=True)
    ind_files = sorted(glob.glob(os.path.join(predpath, PRE_BASE_STR + ".*")))
    print("Find: %s %d %s" % (predpath, len(ind_files), PRE_BASE_STR))                    # :1236
    if not ind_files:
        raise capi.DeepModError("no %s.<chr> index files under %s" % (PRE_BASE_STR, predpath))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    written = []
    with capi.Context(checkpoint.random_model(0), device=local) as ctx:                   # the model is not used in this phase
        for cif in ind_files:
            chrom = cif.split(PRE_BASE_STR)[-1][1:]                                         # :1242
            entries = {s: read_file_list(cif, s) for s in "+-"}
            base_out = entries["+"][1] or entries["-"][1] or (predpath + "/")
            # the sum does not depend on the order: group the reads by detail file, open every file once
            by_file = defaultdict(list)
            for s in "+-":
                for lsp in entries[s][0]:
                    by_file[base_out + "/" + lsp[5]].append((s, lsp[3]))
            clen = 0
            for path in sorted(by_file):
                clen = max(clen, _Detail(path).contig_len([key for _, key in by_file[path]]))
            ctx.set_genome([clen], moptions["Base"])
            for path in sorted(by_file):                   # one file in memory at a time, one accumulate call per strand
                det = _Detail(path)
                for s in "+-":
                    recs = []
                    for st, key in by_file[path]:
                        if st != s:
                            continue
                        rec = det.get(key)
                        if not (rec[4] == chrom and rec[5] == s):                            # :1055-1056
                            print("ERRoR not the same chr (real=%s vs expect=%s) and strand (real=%s VS expect=%s)" % (rec[4], chrom, rec[5], s))
                        recs.append(rec)
                    if recs:
                        ctx.accumulate_records(0, s, np.concatenate([r[0] for r in recs]), np.concatenate([r[1] for r in recs]),
                                               np.concatenate([r[2] for r in recs]), np.concatenate([r[3] for r in recs]))
            for s in "+-":
                path = "%s/mod_pos.%s%s.%s.bed" % (out_dir, chrom, s, moptions["Base"])   # :1043
                if ctx.write_bed(0, s, chrom, path) > 0:
                    written.append(path)
    with open(out_dir + ".done", "a"):
        os.utime(out_dir + ".done", None)
    return {"beds": written}


def convert_run(run_dir):
    """``<outFolder><FileID>`` written with ``--saveDetail 1`` -> the reference's on-disk names: every
    ``<ct>/rnn.pred.detail.dmpd.<batch>`` becomes ``<ct>/rnn.pred.detail.fast5.<batch>`` (HDF5, needs h5py) and the index
    files (per batch and merged) point at the new names.  -> list of HDF5 files written."""
    written = []
    for path in sorted(glob.glob(os.path.join(run_dir, "*", DETAIL_NAME + ".*"))):
        out = path.replace(DETAIL_NAME + ".", "rnn.pred.detail.fast5.")
        to_hdf5(path, out)
        written.append(out)
    for ind in glob.glob(os.path.join(run_dir, "*", "*." + PRE_BASE_STR + ".*")) + glob.glob(os.path.join(run_dir, PRE_BASE_STR + ".*")):
        with open(ind) as fh:
            text = fh.read()
        with open(ind, "w") as fh:
            fh.write(text.replace("/" + DETAIL_NAME + ".", "/rnn.pred.detail.fast5."))
    return written


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 3 and sys.argv[1] == "to-hdf5":
        for p in convert_run(sys.argv[2]):
            print(p)
    else:
        print("usage: python -m deepmod_b200.predetail to-hdf5 <outFolder><FileID>")
        sys.exit(2)
