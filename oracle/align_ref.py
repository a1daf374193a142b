"""CPU oracle for the SAM/CIGAR walk that builds ``base_map_info`` (test infrastructure only).

Restates ``bin/DeepMod_scripts/myDetect.py``: ``handle_line`` (:929-943, best-MAPQ record per read) and the
alignment part of ``handle_record`` (:488-705): clip removal, CIGAR expansion, first/last-match trimming,
strand flip + complement (:661-666), the CpG gap swap (:680-700) and the 'Less Event' test (:702-705).

``run_reference_handle_record`` drives the UNMODIFIED ``handle_line`` + ``handle_record`` in the build container
(tensorflow / h5py stubbed, ``np.int`` shimmed, the per-read HDF5 detail captured in memory), so the restatement
and the GPU path are pinned against the reference's own code.
"""
import re
import sys
import types
from collections import defaultdict

import numpy as np

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a", "N": "N", "n": "n"}   # myCom.py:14-24
NUM = re.compile(r"\d+")
OPS = re.compile(r"[MIDNSHPX=]{1}")

ST_OK, ST_NO_MATCH, ST_LESS_EVENT, ST_FILTERED = 0, 4, 3, 5


def handle_line(line, f5align):
    """myDetect.py:929-943 -> (qname, status)."""
    lsp = line.split("\t")
    qname, flag, rname, pos, mapq, cigar, _, _, _, seq, _ = lsp[:11]
    status = ""
    if qname == "*": status = "qname is *"
    elif int(mapq) == 255: status = "mapq is 255"
    elif int(pos) == 0: status = "pos is 0"
    elif cigar == "*": status = "cigar is *"
    elif rname == "*": status = "rname is *"
    if status != "":
        return qname, status
    if (qname not in f5align) or f5align[qname][0] < int(mapq):
        f5align[qname] = (int(mapq), int(flag), rname, int(pos), cigar, seq)
    return qname, status


def walk(rec, refseq, n_events):
    """Alignment part of handle_record for one record (mapq, flag, rname, pos, cigar, readseq).

    -> dict(status, strand, start_clip, end_clip [read orientation], refbase, readbase, refpos [lists, read
       orientation], first_match_pos, numinsert, numdel, nummismatch)
    """
    mapq, flag, rname, pos, cigar, readseq = rec
    pos = pos - 1                                                     # :519
    strand = "-" if flag & 0x10 else "+"
    numinfo = [int(x) for x in NUM.findall(cigar)]
    mdiinfo = OPS.findall(cigar)
    leftclip = rightclip = 0
    while mdiinfo[0] in "IDNSHPX":                                    # :527-534
        if mdiinfo[0] in "ISX":
            leftclip += numinfo[0]; readseq = readseq[numinfo[0]:]
        if mdiinfo[0] == "H": leftclip += numinfo[0]
        if mdiinfo[0] in "DNX": pos += numinfo[0]
        numinfo = numinfo[1:]; mdiinfo = mdiinfo[1:]
    while mdiinfo[-1] in "IDNSHPX":                                   # :536-540
        if mdiinfo[-1] in "ISX":
            rightclip += numinfo[-1]; readseq = readseq[:-numinfo[-1]]
        if mdiinfo[-1] == "H": rightclip += numinfo[-1]
        numinfo = numinfo[:-1]; mdiinfo = mdiinfo[:-1]
    n_ev = n_events - leftclip - rightclip                            # len(m_event), :541-546
    bmi = []
    firstmatch = lastmatch = first_al = last_al = first_pos = last_pos = None
    nummismatch = numinsert = numdel = 0
    read_ind = 0
    for n, op in zip(numinfo, mdiinfo):                               # :565-621
        for _ in range(n):
            if op == "M":
                bmi.append((refseq[pos], readseq[read_ind], pos))
                if refseq[pos] == readseq[read_ind]:
                    if firstmatch is None: firstmatch = read_ind
                    if lastmatch is None or lastmatch < read_ind: lastmatch = read_ind
                    if first_al is None: first_al = len(bmi) - 1
                    if last_al is None or last_al < len(bmi): last_al = len(bmi) - 1
                    if first_pos is None: first_pos = pos
                    if last_pos is None or last_pos < pos: last_pos = pos
                else:
                    nummismatch += 1
                pos += 1; read_ind += 1
            elif op == "I":
                bmi.append(("-", readseq[read_ind], pos)); read_ind += 1; numinsert += 1
            elif op == "D":
                bmi.append((refseq[pos], "-", pos)); pos += 1; numdel += 1
            elif op == "N":
                bmi.append((refseq[pos], "-", pos)); pos += 1
            elif op == "S":
                read_ind += 1
            elif op == "=":
                bmi.append((refseq[pos], readseq[read_ind], pos))
                if first_pos is None: first_pos = pos
                if last_pos is None or last_pos < pos: last_pos = pos
                pos += 1; read_ind += 1
                if firstmatch is None: firstmatch = read_ind - 1
                if lastmatch is None or lastmatch < read_ind - 1: lastmatch = read_ind - 1
                if last_al is None or last_al < len(bmi): last_al = len(bmi) - 1
                if first_al is None: first_al = len(bmi) - 1
            elif op == "X":
                bmi.append((refseq[pos], readseq[read_ind], pos)); pos += 1; read_ind += 1; nummismatch += 1
    out = dict(strand=strand, rname=rname)
    if firstmatch is None or lastmatch is None:                       # :622-627
        out["status"] = ST_NO_MATCH
        return out
    if strand == "+":                                                 # :630-635
        leftclip += firstmatch
        if n_ev - lastmatch > 1: rightclip += n_ev - lastmatch - 1
    else:
        rightclip += firstmatch
        if n_ev - lastmatch > 1: leftclip += n_ev - lastmatch - 1
    if strand == "+":                                                 # :637-643 (length of the trimmed m_event)
        if n_ev - lastmatch > 1: n_ev2 = (lastmatch + 1) - firstmatch
        elif firstmatch > 0: n_ev2 = n_ev - firstmatch
        else: n_ev2 = n_ev
    else:
        if firstmatch > 0: n_ev2 = (n_ev - firstmatch) - (n_ev - 1 - lastmatch)
        elif n_ev - lastmatch > 1: n_ev2 = n_ev - (n_ev - 1 - lastmatch)
        else: n_ev2 = n_ev
    if firstmatch > 0 or len(bmi) - last_al > 1:                      # :645-657
        if len(bmi) - last_al > 1:
            bmi = bmi[first_al:(last_al + 1 - len(bmi))]
        elif first_al > 0:
            bmi = bmi[first_al:]
    refb = [b[0] for b in bmi]; readb = [b[1] for b in bmi]; refp = [b[2] for b in bmi]
    if strand == "-":                                                 # :661-666
        refb = [COMP.get(x, x) for x in refb[::-1]]
        readb = [COMP.get(x, x) for x in readb[::-1]]
        refp = refp[::-1]
        leftclip, rightclip = rightclip, leftclip
    n = len(refb)
    for ali in range(n):                                              # :680-700 CpG gap swap
        if refb[ali] == "C" and readb[ali] == "C":
            if ali + 1 < n and readb[ali + 1] == "-" and refb[ali + 1] == "G":
                add = 2
                while ali + add < n:
                    if readb[ali + add] == "-" and refb[ali + add] == "G": add += 1
                    else: break
                if ali + add < n and readb[ali + add] == "G" and refb[ali + add] == "G":
                    readb[ali + 1], readb[ali + add] = readb[ali + add], readb[ali + 1]
        if refb[ali] == "G" and readb[ali] == "G":
            if ali - 1 > -1 and readb[ali - 1] == "-" and refb[ali - 1] == "C":
                add = 2
                while ali - add > -1:
                    if readb[ali - add] == "-" and refb[ali - add] == "C": add += 1
                    else: break
                if ali - add > -1 and readb[ali - add] == "C" and refb[ali - add] == "C":
                    readb[ali - 1], readb[ali - add] = readb[ali - add], readb[ali - 1]
    out.update(status=ST_LESS_EVENT if n_ev2 < 50 else ST_OK, start_clip=leftclip, end_clip=rightclip, refbase=refb,
               readbase=readb, refpos=refp, first_match_pos=first_pos, numinsert=numinsert, numdel=numdel,
               nummismatch=nummismatch, n_mapped_events=n_ev2)
    return out


# ---------------------------------------------------------------------------------------------------
# the unmodified reference (build container only)

class _Group(dict):
    def __init__(self):
        super().__init__()
        self.attrs = {}
        self.datasets = {}

    def create_group(self, name):
        g = _Group()
        self[name] = g
        return g

    def create_dataset(self, name, data=None, compression=None):
        self.datasets[name] = np.array(data)


class _File(_Group):
    store = {}

    def __init__(self, path, mode="r"):
        super().__init__()
        self.path = path
        if path in _File.store:
            self.update(_File.store[path])
        _File.store[path] = self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        _File.store[self.path] = dict(self)
        return False

    def flush(self):
        pass

    def close(self):
        pass


def run_reference_handle_record(sess, sam_lines, reads, genome, tmpdir):
    """reads: {qname: dict(ev_mean, ev_stdv, ev_len, ev_base [str list])}; genome: {rname: str}.
    -> {qname: dict(status | predetail arrays + attrs)}"""
    from . import ref_harness
    md = ref_harness.import_myDetect()
    if not hasattr(np, "int"):
        np.int = int                                                  # removed in numpy 1.24 (myDetect.py:660, :752)
    h5 = sys.modules["h5py"]
    h5.File = _File
    md.h5py = h5
    _File.store = {}
    moptions = {"fnum": 7, "hidden": 100, "windowsize": 21, "outLevel": 3, "ConUnk": True, "region": [[None, None, None]],
                "wrkBase": "/w", "outFolder": tmpdir + "/", "FileID": "x"}
    sp_options = defaultdict()
    sp_options["Error"] = defaultdict(list)
    sp_options["rnn"] = (sess, sess.X, sess.Y, sess.init_l, sess.mfpred)
    sp_options["ctfolder"] = tmpdir + "/x/0"
    sp_options["batchid"] = 0
    sp_options["Mod"] = []
    f5data = {}
    for q, rd in reads.items():
        ev = np.zeros(len(rd["ev_mean"]), dtype=ref_harness.EVENT_DTYPE)
        ev["mean"], ev["stdv"] = rd["ev_mean"], rd["ev_stdv"]
        ev["length"] = np.asarray(rd["ev_len"]).astype(np.uint64)
        ev["model_state"] = ["NN" + b + "NN" for b in rd["ev_base"]]
        f5data[q] = ("".join(rd["ev_base"]), ev, None, "/w/%s.fast5" % q)
    sp_param = defaultdict()
    sp_param["f5data"] = f5data
    sp_param["ref_info"] = defaultdict()
    for k, v in genome.items():
        sp_param["ref_info"][k] = v
    f5align = defaultdict()
    old = sys.stdout
    sys.stdout = open("/dev/null", "w")
    try:
        for line in sam_lines:
            if not line or line[0] == "@":
                continue
            sp_param["f5status"] = ""
            sp_param["line"] = line
            md.handle_line(moptions, sp_param, f5align)
        sp_param["f5status"] = ""
        import os
        os.makedirs(sp_options["ctfolder"], exist_ok=True)
        md.handle_record(moptions, sp_options, sp_param, f5align, f5data)
    finally:
        sys.stdout.close()
        sys.stdout = old
    out = {}
    keys = list(f5align.keys())
    store = _File.store.get(sp_options["ctfolder"] + "/rnn.pred.detail.fast5.0", {})
    preds = store.get("pred", {}) if isinstance(store, dict) else {}
    for i, q in enumerate(keys):
        g = preds.get("pred_%d" % i)
        if g is None:
            out[q] = {"written": False}
            continue
        d = g.datasets["predetail"]
        out[q] = {"written": True, "refbase": [x.decode() for x in d["refbase"]], "readbase": [x.decode() for x in d["readbase"]],
                  "refpos": [int(x) for x in d["refbasei"]], "mod_pred": [int(x) for x in d["mod_pred"]], "attrs": dict(g.attrs)}
    out["__errors__"] = {k: list(v) for k, v in sp_options["Error"].items()}
    out["__order__"] = keys
    return out
