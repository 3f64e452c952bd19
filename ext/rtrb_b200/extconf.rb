# Builds the thin Ruby shim over librtrb_b200.so, following the reference's own extension layout
# (ext/fast_4d_matrix/extconf.rb: mkmf + create_makefile(<name>), built by Rake::ExtensionTask into lib/).
require 'mkmf'

extension_name = 'rtrb_b200'
root = File.expand_path('../..', __dir__)
append_cflags('-std=c99')
append_cflags('-O2')
$INCFLAGS << " -I#{File.join(root, 'include')}"
libdir = File.join(root, 'raytracing_rb_b200', 'csrc')
$LDFLAGS << " -L#{libdir} -Wl,-rpath,#{libdir}"
abort 'librtrb_b200.so not found: run `make -C raytracing_rb_b200/csrc` first' unless have_library('rtrb_b200', 'rtrb_abi_version')

dir_config(extension_name)
create_makefile(extension_name)
