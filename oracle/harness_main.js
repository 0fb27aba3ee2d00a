/*
 * harness_main.js — TEST INFRASTRUCTURE.  Drives the reference's register()
 * (module.exports of lib/register.js, evaluated just before this file) with a
 * fake ZooKeeper client that records every call.  zkplus (not vendored) would
 * JSON.stringify the object passed to create()/put(); this ES3 engine has no
 * JSON object, so __stringify below restates ECMA-262 JSON.stringify for the
 * value kinds that occur (objects in property-creation order, arrays, strings,
 * finite numbers, booleans, null; undefined members dropped).
 */
var __register = module.exports.register;
var __unregister = module.exports.unregister;

function __quote(s) {
    var out = '"', i, c, h;
    for (i = 0; i < s.length; i++) {
        c = s.charCodeAt(i);
        if (c === 0x22) out += '\\"';
        else if (c === 0x5c) out += '\\\\';
        else if (c === 0x08) out += '\\b';
        else if (c === 0x0c) out += '\\f';
        else if (c === 0x0a) out += '\\n';
        else if (c === 0x0d) out += '\\r';
        else if (c === 0x09) out += '\\t';
        else if (c < 0x20) { h = c.toString(16); out += '\\u00' + (h.length < 2 ? '0' : '') + h; }
        else out += s.charAt(i);
    }
    return (out + '"');
}

function __stringify(v) {
    var t = typeof (v), parts, k, s, i;
    if (v === null) return ('null');
    if (t === 'string') return (__quote(v));
    if (t === 'number') return (isFinite(v) ? String(v) : 'null');
    if (t === 'boolean') return (String(v));
    if (t === 'undefined' || t === 'function') return (undefined);
    if (Array.isArray(v)) {
        parts = [];
        for (i = 0; i < v.length; i++) { s = __stringify(v[i]); parts.push(s === undefined ? 'null' : s); }
        return ('[' + parts.join(',') + ']');
    }
    parts = [];
    for (k in v) {
        if (!v.hasOwnProperty(k)) continue;
        s = __stringify(v[k]);
        if (s !== undefined) parts.push(__quote(k) + ':' + s);
    }
    return ('{' + parts.join(',') + '}');
}

var __quiet = false;
function __emit(arr) { if (!__quiet) print(__stringify(arr)); }

var __log = { debug: function () {}, info: function () {}, error: function () {}, warn: function () {},
    fatal: function () {}, trace: function () {}, child: function () { return (this); } };

/*
 * ZooKeeper replies arrive on later turns of the Node event loop, never inside the
 * call that issued the request; model that with a FIFO of deferred callbacks that
 * the driver pumps after each synchronous step (the payload is serialised at call
 * time, as a client library does).
 */
var __queue = [];
function __defer(fn) { __queue.push(fn); }
function __pump() {
    while (__queue.length || __timers.length) {
        if (__queue.length)
            (__queue.shift())();
        else
            (__timers.shift())();
    }
}

var __zk = {
    unlink: function (n, cb) { __emit([ 'unlink', n ]); __defer(function () { cb(); }); },
    mkdirp: function (p, cb) { __emit([ 'mkdirp', p ]); __defer(function () { cb(); }); },
    create: function (n, obj, opts, cb) {
        __emit([ 'create', n, __stringify(obj), opts.flags.join('+') ]);
        __defer(function () { cb(); });
    },
    put: function (p, obj, cb) { __emit([ 'put', p, __stringify(obj) ]); __defer(function () { cb(); }); },
    stat: function (p, cb) { __emit([ 'stat', p ]); __defer(function () { cb(null, {}); }); }
};

function __one(line) {
    var cfg = eval('(' + line + ')');
    __hostname = cfg.hostname;
    delete cfg.hostname;
    cfg.log = __log;
    cfg.zk = __zk;
    var result = null;
    try {
        __register(cfg, function (err, znodes) { result = err ? [ 'error', String(err) ] : [ 'registered' ].concat(znodes); });
        __pump();
    } catch (e) {
        result = [ 'throw', String(e.message || e) ];
    }
    __emit(result === null ? [ 'pending' ] : result);
}

function harness_run(input, repeat) {
    var lines = input.split('\n'), i, r, t0, t1, n = 0;
    if (repeat > 0) {
        __quiet = true;
        t0 = new Date().getTime();
        for (r = 0; r < repeat; r++)
            for (i = 0; i < lines.length; i++)
                if (lines[i].length) { __one(lines[i]); n++; }
        t1 = new Date().getTime();
        print('{"records":' + n + ',"ms":' + (t1 - t0) + '}');
        return;
    }
    for (i = 0; i < lines.length; i++)
        if (lines[i].length)
            __one(lines[i]);
}
