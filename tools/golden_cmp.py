import sys, numpy as np, torch
sys.path.insert(0, ".")
import astc_encoder_b200 as A
from PIL import Image
img = np.array(Image.open("tests/golden/leaf.png").convert("RGBA"))[::-1].copy()
gold = np.frombuffer(open("tests/golden/leaf.astc", "rb").read()[16:], dtype=np.uint8).reshape(-1, 16)
out = A.encode_astc(torch.from_numpy(img).cuda(), A.encode_option(has_alpha=True)).cpu().numpy()
same = int((out == gold).all(axis=1).sum())
print(f"identical to golden: {same} of {len(gold)} = {same / len(gold) * 100:.3f} %")
