"""Train / test driver — the orchestration of src/train.lua over the `Model` mirror:

    python -m aocr.train -phase train -data_base_dir D -data_path train.txt -val_data_path val.txt -model_dir M ...

`train()` follows src/train.lua:68-216 statement by statement (per-step perplexity log, checkpoint cadence with the
`final-model` hand-over, validation pass, learning-rate decay on a validation-loss increase, end-of-epoch checkpoint and
validation); `main()` follows :218-296 (logger, model create / load, data sets, dictionary).  The options are the
reference's (`cmd:option`, train.lua:18-64) with the same names and defaults; `-gpu_id` is 1-based as there.
The hot path behind `model.step` is libaocr.so; nothing here computes.
"""
import argparse
import math
import os
import shutil
import sys
import time

from .data import DataGen
from .model import Model


class Logger:
    """src/utils/logging.lua: every message to the log file and to stdout, time-stamped"""

    def __init__(self, log_path):
        self.f = open(log_path, "a") if log_path else None

    def info(self, msg):
        line = "[%s] %s" % (time.strftime("%Y-%m-%d %H:%M:%S"), msg)
        print(line, flush=True)
        if self.f:
            self.f.write(line + "\n")
            self.f.flush()

    def shutdown(self):
        if self.f:
            self.f.close()
            self.f = None


def _exp(x):
    try:
        return math.exp(x)
    except (OverflowError, ZeroDivisionError):
        return float("inf")


def _validate(model, val_data, batch_size, num_batches_val, beam_size, trie, logging):
    """train.lua:136-161 (and the identical block at :181-205)"""
    val_loss, val_num_samples, val_num_nonzeros, val_accuracy = 0.0, 0, 0, 0.0
    b = 1
    while b <= num_batches_val:
        if b % 100 == 0:
            logging.info("%d" % b)
        val_batch = val_data.nextBatch(batch_size)
        if val_batch is None:
            val_data.shuffle()
            if num_batches_val >= math.inf:
                break
        else:
            real_batch_size = val_batch[0].shape[0]
            b += 1
            step_loss, stats = model.step(val_batch, True, beam_size, trie)
            val_loss += step_loss
            val_num_samples += real_batch_size
            val_num_nonzeros += stats[0]
            val_accuracy += stats[1]
    return val_loss, val_num_samples, val_num_nonzeros, val_accuracy


def train(model, phase, batch_size, num_epochs, train_data, val_data, model_dir, steps_per_checkpoint, num_batches_val,
          beam_size, visualize, output_dir, trie, opt, logging):
    """src/train.lua:68-216"""
    loss, num_seen, num_samples, num_nonzeros, accuracy = 0.0, 0, 0, 0, 0.0
    if phase == "train":
        forward_only = False
    elif phase == "test":
        if visualize:
            model.vis(output_dir)
        forward_only = True
        num_epochs = 1
        model.global_step = 0
    else:
        raise AssertionError("phase must be either train or test")
    learning_rate = model.optim_state.get("learningRate") or opt.learning_rate          # :86-89
    learning_rate = max(learning_rate, opt.learning_rate_min)
    model.optim_state["learningRate"] = learning_rate
    logging.info("Lr: %f" % learning_rate)
    prev_val_loss = None

    def decay(val_loss):                                                                # :163-168 / :207-212
        nonlocal prev_val_loss
        lr = model.optim_state["learningRate"]
        if prev_val_loss is not None and val_loss > prev_val_loss and lr > opt.learning_rate_min:
            lr = max(lr * opt.lr_decay, opt.learning_rate_min)
            model.optim_state["learningRate"] = lr
            logging.info("Decay lr, current Lr: %f" % lr)
        prev_val_loss = val_loss

    for epoch in range(1, num_epochs + 1):
        if not forward_only:
            train_data.shuffle()
        while True:
            train_batch = train_data.nextBatch(batch_size)
            if train_batch is None:
                break
            real_batch_size = train_batch[0].shape[0]
            step_loss, stats = model.step(train_batch, forward_only, beam_size, trie)
            # :103 logs the perplexity of the totals BEFORE this step (NaN on the first)
            logging.info("%f" % (_exp(loss / num_nonzeros) if num_nonzeros else float("nan")))
            num_seen += 1
            num_samples += real_batch_size
            num_nonzeros += stats[0]
            if forward_only:
                accuracy += stats[1]
            else:
                loss += step_loss
            model.global_step += 1
            if model.global_step % steps_per_checkpoint == 0:
                if forward_only:
                    logging.info("Number of samples %d - Accuracy = %f" % (num_samples, accuracy / num_samples))
                else:
                    logging.info("Step %d - training perplexity = %f" % (model.global_step, _exp(loss / num_nonzeros)))
                    logging.info("Saving model")
                    model_path = os.path.join(model_dir, "model-%d" % model.global_step)
                    final_tmp = os.path.join(model_dir, ".final-model.tmp")
                    final_model_path = os.path.join(model_dir, "final-model")
                    saved = model.save(model_path)                                    # :125
                    logging.info("Model saved to %s" % model_path)
                    shutil.copyfile(saved, final_tmp)                                 # :127-128: cp, then atomic mv
                    os.replace(final_tmp, final_model_path + os.path.splitext(saved)[1])
                    num_seen, num_nonzeros, loss, accuracy = 0, 0, 0.0, 0.0
                    logging.info("Evaluating model on %s batches of validation data" % num_batches_val)
                    vl, vs, vn, va = _validate(model, val_data, batch_size, num_batches_val, beam_size, trie, logging)
                    logging.info("Step %d - Val Accuracy = %f, loss = %f" % (model.global_step, va / max(vs, 1), _exp(vl / max(vn, 1))))
                    decay(vl)
        if forward_only:
            logging.info("Epoch: %d Number of samples %d - Accuracy = %f" % (epoch, num_samples, accuracy / max(num_samples, 1)))
        else:
            model_path = os.path.join(model_dir, "model-%d" % model.global_step)
            model.save(model_path)                                                    # :176-178
            logging.info("Model saved to %s" % model_path)
            logging.info("Evaluating model on %s batches of validation data" % num_batches_val)
            vl, vs, vn, va = _validate(model, val_data, batch_size, num_batches_val, beam_size, trie, logging)
            logging.info("Epoch: %d, Step %d - Val Accuracy = %f, loss = %f" % (epoch, model.global_step, va / max(vs, 1),
                                                                               _exp(vl / max(vn, 1))))
            decay(vl)
    return {"num_samples": num_samples, "accuracy": accuracy, "loss": loss, "num_nonzeros": num_nonzeros}


def build_parser():
    """cmd:option list of src/train.lua:18-64, same names and defaults"""
    ap = argparse.ArgumentParser(prog="aocr.train", prefix_chars="-", allow_abbrev=False)
    a = ap.add_argument
    a("-data_base_dir", default="/n/rush_lab/data/image_data/90kDICT32px")
    a("-data_path", default="/n/rush_lab/data/image_data/train_shuffled_shuffled_words.txt")
    a("-val_data_path", default="/n/rush_lab/data/image_data/val_shuffled_words.txt")
    a("-model_dir", default="train")
    a("-log_path", default="log.txt")
    a("-output_dir", default="results")
    a("-steps_per_checkpoint", type=int, default=1000)
    a("-num_batches_val", type=float, default=math.inf)
    a("-beam_size", type=int, default=1)
    a("-use_dictionary", action="store_true")
    a("-allow_digit_prefix", action="store_true")
    a("-dictionary_path", default="/n/rush_lab/data/image_data/train_dictionary.txt")
    a("-num_epochs", type=int, default=1000)
    a("-batch_size", type=int, default=400)
    a("-learning_rate", type=float, default=0.1)
    a("-learning_rate_min", type=float, default=0.001)
    a("-lr_decay", type=float, default=0.5)
    a("-dropout", type=float, default=0.0)
    a("-target_embedding_size", type=int, default=20)
    a("-input_feed", action="store_true")
    a("-encoder_num_hidden", type=int, default=512)
    a("-encoder_num_layers", type=int, default=1)
    a("-decoder_num_layers", type=int, default=2)
    a("-target_vocab_size", type=int, default=26 + 10 + 3)
    a("-phase", default="test")
    a("-gpu_id", type=int, default=1)
    a("-load_model", action="store_true")
    a("-visualize", action="store_true")
    a("-seed", type=int, default=910820)
    a("-max_decoder_l", type=int, default=50)
    a("-max_encoder_l", type=int, default=80)
    a("-prealloc", action="store_true")
    # not in the reference: its data layer forces every image to width 100 (data_gen.lua:78); -1 lifts that
    a("-fixed_width", type=int, default=100)
    return ap


def main(argv=None):
    """src/train.lua:218-296"""
    opt = build_parser().parse_args(argv)
    logging = Logger(opt.log_path)
    logging.info("Command Line Arguments:")
    logging.info(" ".join(sys.argv[1:] if argv is None else argv))
    logging.info("End Command Line Arguments")
    assert opt.gpu_id > 0, "the hot path has no CPU implementation (libaocr is sm_100a only): -gpu_id must be >= 1"
    logging.info("Using CUDA on GPU %d" % opt.gpu_id)
    logging.info("Building model")
    model = Model(log=logging.info, device=opt.gpu_id - 1)
    final_model = os.path.join(opt.model_dir, "final-model")
    if opt.load_model and (os.path.isfile(final_model) or os.path.isfile(final_model + ".npz")):
        logging.info("Loading model from %s" % final_model)
        model.load(final_model, vars(opt))
    else:
        logging.info("Creating model with fresh parameters")
        model.create(vars(opt))
    os.makedirs(opt.model_dir, exist_ok=True)
    if opt.visualize:
        os.makedirs(opt.output_dir, exist_ok=True)
    fixed = opt.fixed_width if opt.fixed_width > 0 else None
    logging.info("Data base dir %s" % opt.data_base_dir)
    logging.info("Load training data from %s" % opt.data_path)
    from .capi import PinnedRing
    prefetch = 2
    train_data = DataGen(opt.data_base_dir, opt.data_path, 10.0, fixed_width=fixed, log=logging.info, seed=opt.seed,
                         prefetch=prefetch, alloc=PinnedRing(prefetch + 4))     # page-locked batches, built ahead
    logging.info("Training data loaded from %s" % opt.data_path)
    val_data = None
    if opt.phase == "train":
        logging.info("Load validation data from %s" % opt.val_data_path)
        val_data = DataGen(opt.data_base_dir, opt.val_data_path, 10.0, fixed_width=fixed, log=logging.info, seed=opt.seed)
        logging.info("Validation data loaded from %s" % opt.val_data_path)
    trie = None
    if opt.use_dictionary:
        from .capi import Trie
        logging.info("Load dictionary from %s" % opt.dictionary_path)
        trie = Trie(path=opt.dictionary_path, allow_digit_prefix=opt.allow_digit_prefix)
    train(model, opt.phase, opt.batch_size, opt.num_epochs, train_data, val_data, opt.model_dir, opt.steps_per_checkpoint,
          opt.num_batches_val, opt.beam_size, opt.visualize, opt.output_dir, trie, opt, logging)
    logging.shutdown()
    model.shutdown()


if __name__ == "__main__":
    main()
