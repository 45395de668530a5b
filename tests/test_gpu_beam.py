"""§8f-1: beam search (beam_size > 1) and dictionary-constrained decode (src/model/model.lua:380-387,405-445,460-536,
573-585; src/utils/utils.lua:177-218) through aocr_decode_beam, against the oracle's step-by-step restatement.
Token parity is exact outside ties: a row is a tie when the margin between the last continuation a step takes and the
best one it leaves out (or between the two best final beams) is below 1e-4 - the reference's `topk` leaves the order
of equal scores unspecified."""
import numpy as np
import pytest

from oracle import Config, Oracle, init_params, init_bn_stats, make_batch, load_dictionary, flatten_trie
from oracle.synth import str2numlist
from parity_util import make_handle, rel_err, TOL

pytestmark = pytest.mark.gpu
TIE = 1e-4


def _run(cfg, batch, beam, words=None, allow_digit_prefix=False, seed=910820, duplicates=False):
    from aocr import Trie
    params, bn = init_params(cfg, seed), init_bn_stats(cfg)
    orc = Oracle(cfg, params, bn)
    otrie = load_dictionary(words, allow_digit_prefix) if words is not None else None
    o = orc.decode_beam(batch["images"], batch["targets"], batch["targets_eval"], beam, trie=otrie)
    h = make_handle(cfg, params, bn)
    trie = Trie(words=words, allow_digit_prefix=allow_digit_prefix) if words is not None else None
    if trie is not None:                                    # the C restatement of loadDictionary = the oracle's
        assert np.array_equal(trie.numpy(), flatten_trie(otrie))
    g = h.decode_beam(batch["images"], batch["targets"], batch["targets_eval"], beam, trie)
    h.close()
    B = batch["images"].shape[0]
    gaps, fgap = o["tie_gaps"], o["final_gap"]
    if duplicates:      # beams padded with copies of one hypothesis (model.lua:424-436): exact-zero margins between copies
        gaps, fgap = np.where(gaps == 0, np.inf, gaps), np.where(fgap == 0, np.inf, fgap)
    tie = (gaps.min(axis=1) < TIE) | (fgap < TIE)
    ok = ~tie
    assert ok.sum() > 0, "every row is a tie: the case checks nothing"
    assert np.array_equal(g["labels"][ok], o["labels"][ok]), (g["labels"][ok], o["labels"][ok])
    assert rel_err(g["pred_scores"][ok], o["pred_scores"][ok]) < TOL
    assert rel_err(g["gold_scores"], o["gold_scores"]) < TOL
    assert abs(g["loss_sum"] - o["loss_sum"]) < TOL * abs(o["loss_sum"])
    if tie.sum() == 0:
        assert g["num_correct"] == o["num_correct"]
    return g, o, tie


@pytest.mark.parametrize("beam", [2, 5])
def test_beam_search_matches_oracle(beam):
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=10)
    _run(cfg, make_batch(4, 100, 6, seed=81), beam)


def test_beam_one_equals_greedy():
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=10)
    batch = make_batch(4, 100, 6, seed=82)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    h = make_handle(cfg, params, bn)
    a = h.decode_greedy(batch["images"], batch["targets"], batch["targets_eval"])
    b = h.decode_beam(batch["images"], batch["targets"], batch["targets_eval"], 1)
    h.close()
    assert np.array_equal(a["labels"], b["labels"]) and a["num_correct"] == b["num_correct"]
    assert rel_err(b["pred_scores"], a["pred_scores"]) < 1e-6 and rel_err(b["gold_scores"], a["gold_scores"]) < 1e-6


def test_beam_wider_than_the_state_chunks_the_batch():
    """beam * batch > 2 x batch_size rows: the beam pass walks the batch in chunks; beam clipped to the vocabulary"""
    cfg = Config(batch_size=6, max_encoder_l=30, max_decoder_l=6)
    _run(cfg, make_batch(6, 100, 4, seed=83), 5)             # 12 state rows -> chunks of 2 images
    cfg = Config(batch_size=20, max_encoder_l=30, max_decoder_l=5)
    _run(cfg, make_batch(3, 100, 3, seed=84), 64)            # clipped to 39 beams, one image per chunk


WORDS = ["hello", "help", "held", "h3", "abc", "a", "zz9", "world", "w0rd", "0", "42", "4", "xyz"]


def test_dictionary_constrained_decode_matches_oracle():
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=10)
    g, o, tie = _run(cfg, make_batch(4, 100, 6, seed=85), 3, words=WORDS)
    # every decoded string is a dictionary word (rows outside ties)
    for row in g["labels"][~tie]:
        ids = []
        for v in row.tolist():
            if v == 3:
                break
            ids.append(v)
        s = "".join(chr(v - 14 + 97) if v > 13 else chr(v - 4 + 48) for v in ids if v > 3)
        # PAD is always admissible (model.lua:472), so a hypothesis may stop inside a word: a prefix of a dictionary word
        assert any(w.startswith(s) for w in WORDS), s
        if 3 in row.tolist():
            assert s in WORDS, s                             # EOS only exists at the end of a word


def test_dictionary_with_fewer_first_characters_than_beams_and_digit_prefix():
    """the first step pads the beam with the best admissible character (model.lua:424-436); allow_digit_prefix loops
    the root on digits and EOS (utils.lua:195-201)"""
    cfg = Config(batch_size=3, max_encoder_l=30, max_decoder_l=8)
    _run(cfg, make_batch(3, 100, 5, seed=86), 4, words=["ab", "ac"], duplicates=True)   # one admissible first character, 4 beams
    _run(cfg, make_batch(3, 100, 5, seed=87), 3, words=["cat", "car"], allow_digit_prefix=True)


def test_model_step_routes_beam_and_trie():
    from aocr import Model, Trie
    cfg = Config(batch_size=4, max_encoder_l=30, max_decoder_l=10)
    params, bn = init_params(cfg, 910820), init_bn_stats(cfg)
    m = Model(log=lambda s: None).create(dict(batch_size=4, max_encoder_l=30, max_decoder_l=10, input_feed=True))
    m.set_parameters(params, bn)
    b = make_batch(4, 100, 6, seed=88)
    bl = [b["images"], b["targets"], b["targets_eval"], b["num_nonzeros"], None]
    o = Oracle(cfg, params, bn).decode_beam(b["images"], b["targets"], b["targets_eval"], 3, trie=load_dictionary(WORDS))
    loss, stats = m.step(bl, True, 3, Trie(words=WORDS))
    assert abs(loss - o["loss_sum"]) < TOL * abs(o["loss_sum"]) and stats[0] == b["num_nonzeros"]
    m.shutdown()
