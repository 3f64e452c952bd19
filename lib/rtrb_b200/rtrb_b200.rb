# Stub declarations + loader for the CUDA tracing core shim, same convention as the reference's
# lib/fast_4d_matrix/fast_4d_matrix.rb (stubs first, then `require_relative '../<name>'` at :38).
module Rtrb
  class Renderer
    # method stubs
    def initialize(world); raise NotImplementedError; end
    def render(camera, seed = 1, precision = 1); raise NotImplementedError; end
    def stats; raise NotImplementedError; end
  end
end

require_relative '../rtrb_b200'

# Frame-level drop-in for the reference's pixel loops.  `Camera#render_cuda(file_path)` replaces
# render_sync(file_path) (src/camera.rb:101-110) / render_fork(file_path, n) (:41-68): one call
# returns the finished 8-bit rows, which go through the existing canvas + save_image (:36-39).
module Alex
  class Camera
    def render_cuda(file_path, seed = 1)
      @rtrb ||= Rtrb::Renderer.new(@world)
      rgb = @rtrb.render(self, seed)           # H*W*3 bytes (RGB8), row = y, column = x
      @height.times do |y|
        @width.times do |x|
          r, g, b = rgb.getbyte((y * @width + x) * 3), rgb.getbyte((y * @width + x) * 3 + 1), rgb.getbyte((y * @width + x) * 3 + 2)
          # same canvas coordinates render_sync uses: render_at returns [x, H-1-y] (camera.rb:98,105)
          @canvas.point(x, @height - 1 - y, PNG::Color.new(r, g, b))
        end
      end
      save_image(file_path)
    end
  end
end
