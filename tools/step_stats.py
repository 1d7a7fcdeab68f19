import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import astc_images as I
import __graft_entry__ as g
pkg = g.load_package()
img = I.photo_like(2048, 2048, seed=2024)
ctx = pkg.Context(pkg.config_init(1, 6, 6, 60.0, 32))
for _ in range(2):
    ctx.compress_image(img)
ctx.close()
