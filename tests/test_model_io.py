"""Model interchange: hlaModelToObj / hlaModelFromObj (reference R/HIBAG.R:1041-1178,
src/HIBAG.cpp:817-958), the on-disk forms, and the RDX2 reader that loads published R workspaces.
Host-only: no GPU needed (the model container does not touch the device)."""
import gzip
import os
import struct

import numpy as np
import pytest

from tests import helpers


def _golden_model(api, n=12):
    ml = helpers.load_golden("modellist_a.npz")
    m = api.HLAModel(len(ml["snp_id"]), len(ml["hla_allele"]), [str(a) for a in ml["hla_allele"]],
                     [str(s) for s in ml["snp_id"]])
    m.n_samp = ml["samp_num"].shape[1]
    for k in range(n):
        c = helpers.golden_classifier(ml, k)
        m.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"], samp_num=c["samp_num"], oob_acc=c["oob_acc"])
    return m, ml


def _same(a, b):
    assert a.num_classifiers() == b.num_classifiers() and a.n_snp == b.n_snp and a.n_hla == b.n_hla
    assert [str(x) for x in a.hla_allele] == [str(x) for x in b.hla_allele]
    for k in range(a.num_classifiers()):
        assert helpers.classifier_diff(a.classifier(k), b.classifier(k)) == "", k
        assert np.array_equal(a.classifier(k)["samp_num"], b.classifier(k)["samp_num"])


def test_to_obj_from_obj_round_trip(built):
    from hibag_b200 import api
    m, ml = _golden_model(api)
    obj = m.to_obj()
    c0 = obj["classifiers"][0]
    # the reference's layout: 1-based SNP indices, allele labels, "0101" strings of n_snp characters
    assert c0["snpidx"].min() >= 1 and set("".join(c0["haplos"]["haplo"])) <= {"0", "1"}
    assert all(len(s) == len(c0["snpidx"]) for s in c0["haplos"]["haplo"])
    assert set(c0["haplos"]["hla"]) <= set(obj["hla_allele"])
    _same(api.HLAModel.from_obj(obj), m)
    assert api.hlaModelFromObj is api.HLAModel.from_obj and api.hlaModelToObj is api.HLAModel.to_obj


@pytest.mark.parametrize("ext", [".json", ".npz"])
def test_save_load_is_bit_exact(built, tmp_path, ext):
    from hibag_b200 import api
    m, _ = _golden_model(api)
    path = str(tmp_path / ("model" + ext))
    m.save(path)
    back = api.HLAModel.load(path)
    _same(back, m)
    assert back.snp_id == m.snp_id and back.n_samp == m.n_samp


def test_classifier_without_bootstrap_counts_and_mismatched_counts(built):
    """classifier() sized its samp_num buffer from the Python-side n_samp; it now asks the
    library, and from_obj rejects a samp_num whose length disagrees with n_samp."""
    from hibag_b200 import api
    m, ml = _golden_model(api, 2)
    obj = m.to_obj()
    obj["n_samp"] = 0                                  # a model object that lost n.samp
    back = api.HLAModel.from_obj(obj)
    assert len(back.classifier(0)["samp_num"]) == ml["samp_num"].shape[1]
    obj2 = m.to_obj()
    obj2["classifiers"][1]["samp_num"] = obj2["classifiers"][1]["samp_num"][:-3]
    with pytest.raises(ValueError):
        api.HLAModel.from_obj(obj2)
    obj3 = m.to_obj()
    for c in obj3["classifiers"]:
        c["samp_num"] = None
    assert len(api.HLAModel.from_obj(obj3).classifier(0)["samp_num"]) == 0


# ---- RDX2: a tiny XDR writer (test-only) so that the reader is covered without /root/reference ----
def _xdr_int(v):
    return struct.pack(">i", v)


def _xdr_chars(s):
    b = s.encode("latin-1")
    return _xdr_int(9 | (1 << 18)) + _xdr_int(len(b)) + b        # CHARSXP, ASCII flag


def _xdr_vec(v, names=None):
    """INTSXP / REALSXP / STRSXP / VECSXP with an optional names attribute"""
    flag_attr = (1 << 9) if names is not None else 0
    if isinstance(v, list) and all(isinstance(x, str) for x in v) and v:
        body = _xdr_int(16 | flag_attr) + _xdr_int(len(v)) + b"".join(_xdr_chars(x) for x in v)
    elif isinstance(v, list):
        body = _xdr_int(19 | flag_attr) + _xdr_int(len(v)) + b"".join(v)
    elif np.asarray(v).dtype.kind == "f":
        a = np.asarray(v, dtype=">f8")
        body = _xdr_int(14 | flag_attr) + _xdr_int(a.size) + a.tobytes()
    else:
        a = np.asarray(v, dtype=">i4")
        body = _xdr_int(13 | flag_attr) + _xdr_int(a.size) + a.tobytes()
    if names is not None:
        # pairlist with one tagged element (names), then NILVALUE
        body += _xdr_int(2 | (1 << 10)) + _xdr_int(1) + _xdr_chars("names") + _xdr_vec(list(names)) + _xdr_int(0xFE)
    return body


def _xdr_list(d):
    return _xdr_vec([v if isinstance(v, bytes) else _xdr_vec(v) for v in d.values()], names=list(d))


def _write_rdata(path, name, payload):
    raw = b"RDX2\nX\n" + _xdr_int(2) + _xdr_int(0x00040300) + _xdr_int(0x00020300)
    raw += _xdr_int(2 | (1 << 10)) + _xdr_int(1) + _xdr_chars(name) + payload + _xdr_int(0xFE)
    with gzip.open(path, "wb") as f:
        f.write(raw)


def test_rdata_workspace_with_a_model_object_loads_without_r(built, tmp_path):
    from hibag_b200 import api
    m, ml = _golden_model(api, 3)
    obj = m.to_obj()
    cls = []
    for c in obj["classifiers"]:
        cls.append(_xdr_list({"samp.num": c["samp_num"], "haplos": _xdr_list({
            "freq": c["haplos"]["freq"], "hla": list(c["haplos"]["hla"]), "haplo": list(c["haplos"]["haplo"])}),
            "snpidx": c["snpidx"], "outofbag.acc": np.array([c["outofbag_acc"]])}))
    payload = _xdr_list({"n.samp": np.array([obj["n_samp"]]), "n.snp": np.array([obj["n_snp"]]),
                         "sample.id": ["S%d" % i for i in range(obj["n_samp"])], "snp.id": list(obj["snp_id"]),
                         "hla.locus": ["A"], "hla.allele": list(obj["hla_allele"]), "classifiers": _xdr_vec(cls)})
    path = str(tmp_path / "model.RData")
    _write_rdata(path, "mobj", payload)
    back = api.hlaModelFromRData(path)
    _same(back, m)
    assert back.hla_locus == "A" and back.snp_id == m.snp_id
    _same(api.HLAModel.load(path), m)


def test_shipped_modellist_rdata_loads_through_the_package(built):
    """The reference's own published-model fixture (inst/extdata/ModelList.RData, xz-compressed RDX2)
    -> HLAModel, equal to the committed golden vectors of the same file."""
    path = "/root/reference/inst/extdata/ModelList.RData"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not on this box")
    from hibag_b200 import api
    m = api.hlaModelFromRData(path, "A")
    ml = helpers.load_golden("modellist_a.npz")
    assert m.num_classifiers() == 100 and [str(a) for a in ml["hla_allele"]] == m.hla_allele
    for k in range(100):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)
    with pytest.raises(ValueError):
        api.hlaModelFromRData(path, "no such locus")
