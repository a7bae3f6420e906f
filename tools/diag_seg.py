import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_torch as R
from vae_segmentation_b200 import joint_model as jm, train_step as ts
from vae_segmentation_b200.synthetic import synth_image, synth_label

for patch, batch in ((32, 2), (64, 1), (64, 2), (48, 1)):
    torch.manual_seed(51)
    sd = R.init_seg_state()
    img = synth_image(batch, patch)
    with torch.no_grad():
        ref = R.seg_forward(sd, img)
    seg = jm.Segmentation(1, 2, norm_type=1); seg.load_state_dict(sd); seg = seg.cuda().set_precision("fp32")
    with torch.no_grad():
        p = seg.predict(img.cuda())
    print("P=%d B=%d plain: %.3e" % (patch, batch, (p.cpu() - ref).abs().max().item()))
    arena = ts.FlatArena(seg)
    with torch.no_grad():
        p = seg.predict(img.cuda())
    print("P=%d B=%d arena: %.3e" % (patch, batch, (p.cpu() - ref).abs().max().item()))
    flat = torch.cat([v.reshape(-1) for v in sd.values()])
    print("   arena data equal:", torch.equal(arena.data.cpu(), flat), [ (k, p.data_ptr() % 16) for k, p in seg.named_parameters() if p.data_ptr() % 16][:4])
