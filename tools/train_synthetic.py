"""End-to-end training on a synthetic in-memory corpus: the reference-shaped callables (build_model,
f_log_probs, pred_probs), train.Trainer as f_grad_shared / f_update, train_loop.fit as the train() bookkeeping.
Writes the reference's files (model_current.npz, model_best_so_far.npz, model_best.npz, train_valid_test.txt,
model_options.pkl) into --out.  Needs a GPU.  (Written at the end of round 1, after the GPU budget was spent: not
run on a GPU yet.)
    python tools/train_synthetic.py --out /tmp/stat_run --epochs 3"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import data_engine, model_attention as ma, train_loop
from video_description_with_spatial_temporal_attention_b200.train import Trainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='/tmp/stat_run')
    ap.add_argument('--epochs', type=int, default=3)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--optimizer', default='adam')
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    o = stat.default_options(dim=a.dim, dim_word=a.dim, ctxg_dim=256, ctxl_dim=512, ctxm_dim=512, n_words=500,
                             global_proj=True)
    model = ma.Attention()
    model.engine = data_engine.synthetic_engine(o, n_videos=48, caps_per_video=4, T=8, R=4, mb_size_train=16,
                                                mb_size_test=16)
    params = model.init_params(o)
    tparams = model.init_tparams(params)
    r = model.build_model(tparams, o)
    use_noise, inps, alphas, cost = r[1], list(r[2:10]), list(r[10:14]), r[14]
    f_log_probs = ma.function(inps, -cost)
    f_alphas = ma.function(inps, alphas)
    trainer = Trainer.from_tparams(tparams, o, optimizer=a.optimizer, alpha_c=0.70602, decay_c=1e-4, clip_c=10.)
    res = train_loop.fit(model, tparams, o, trainer, f_log_probs, f_alphas, a.out + os.sep, use_noise=use_noise,
                         max_epochs=a.epochs, validFreq=6, dispFreq=3, sampleFreq=10 ** 9, patience=10)
    print('train_err %.4f valid_err %.4f test_err %.4f' % res)


if __name__ == '__main__':
    main()
